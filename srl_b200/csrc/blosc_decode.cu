// native decode of the compressed sample-wire payloads, straight into the buffer's pinned
// staging block.  The reference compresses a leaf with the third-party `blosc` package -- blosc.compress(payload,
// typesize=4, cname='lz4'), base/namedarray.py:126,150 -- and decodes it with blosc.decompress into a fresh bytes object per
// leaf per message (namedarray.py:184-185), which the buffer then copies again.  `blosc` is not vendored in the reference and
// is not pinned in its requirements; it is absent from this image.  What is restated here is the PUBLISHED format:
//   * the LZ4 block format (lz4_Block_format.md): sequences of {token, literal length bytes, literals, 2-byte little-endian
//     offset, match length bytes}; the last sequence ends after its literals.  PINNED: tests/test_wire_native.py decodes
//     blocks written by liblz4 itself (pyarrow's `lz4_raw` codec).
//   * the Blosc-1 frame (c-blosc README_HEADER.rst / blosc.c `blosc_d`): 16-byte header {version, versionlz, flags, typesize,
//     nbytes, blocksize, cbytes}; flags bit 0 byte-shuffle, bit 1 memcpyed, bit 2 bit-shuffle, bit 4 "do not split", bits 5-7
//     the codec family (1 = LZ4 / LZ4HC); then one int32 start offset per block; each block is `nsplits` streams of {int32
//     compressed size, bytes}, a stream whose compressed size equals its decoded size being a plain copy; a block is split into
//     `typesize` streams unless bit 4 is set, it is the shorter last block, typesize > 16 or blocksize / typesize < 128;
//     after decoding, a shuffled block is un-shuffled (byte j of element i sits at j * n_elements + i; the bytes behind the
//     last whole element are not shuffled).  UNPINNED against blosc itself (no blosc here to write vectors): the tests build
//     frames with a writer that follows the same description (tests/blosc1_writer.py).  srl_b200/wire.py cross-checks this
//     decoder against the real package once per process wherever that package is importable.
// Pure host code (no CUDA call): usable without a GPU.  Every length and offset is checked; malformed input is an error code,
// never an out-of-bounds access.
#include <string.h>

#include <atomic>
#include <thread>
#include <vector>

#include "common.cuh"

namespace srl {
namespace {

inline uint32_t le32(const unsigned char* p) {
  return static_cast<uint32_t>(p[0]) | (static_cast<uint32_t>(p[1]) << 8) | (static_cast<uint32_t>(p[2]) << 16) |
         (static_cast<uint32_t>(p[3]) << 24);
}

// One LZ4 block.  Returns the number of bytes written, or -1 for malformed input / an output that does not fit.
long long lz4_block_decode(const unsigned char* src, size_t n, unsigned char* dst, size_t cap) {
  const unsigned char* ip = src;
  const unsigned char* const iend = src + n;
  unsigned char* op = dst;
  unsigned char* const oend = dst + cap;
  for (;;) {
    if (ip >= iend) return -1;
    const unsigned token = *ip++;
    size_t lit = token >> 4;
    if (lit == 15) {
      unsigned b;
      do {
        if (ip >= iend) return -1;
        b = *ip++;
        lit += b;
      } while (b == 255);
    }
    if (lit > static_cast<size_t>(iend - ip) || lit > static_cast<size_t>(oend - op)) return -1;
    // short literal runs dominate: one fixed 16-byte copy where both buffers have the room (the surplus is overwritten)
    if (lit <= 16 && iend - ip >= 16 && oend - op >= 16) {
      memcpy(op, ip, 16);
    } else {
      memcpy(op, ip, lit);
    }
    op += lit;
    ip += lit;
    if (ip == iend) break;  // the last sequence: literals only
    if (iend - ip < 2) return -1;
    const size_t off = static_cast<size_t>(ip[0]) | (static_cast<size_t>(ip[1]) << 8);
    ip += 2;
    if (off == 0 || off > static_cast<size_t>(op - dst)) return -1;
    size_t ml = token & 15u;
    if (ml == 15) {
      unsigned b;
      do {
        if (ip >= iend) return -1;
        b = *ip++;
        ml += b;
      } while (b == 255);
    }
    ml += 4;
    if (ml > static_cast<size_t>(oend - op)) return -1;
    const unsigned char* m = op - off;
    if (off >= 16 && static_cast<size_t>(oend - op) >= ml + 16) {
      for (size_t i = 0; i < ml; i += 16) memcpy(op + i, m + i, 16);  // source chunk i ends at or before destination chunk i starts
    } else {
      // a match that overlaps its own output repeats a pattern of period `off`: what lies between m and the write position
      // is whole periods of it, so each copy can take all of that (the copies double in length and never overlap)
      for (size_t done = 0; done < ml;) {
        size_t c = off + done;
        if (c > ml - done) c = ml - done;
        memcpy(op + done, m, c);
        done += c;
      }
    }
    op += ml;
  }
  return op - dst;
}

struct Frame {
  unsigned flags, typesize;
  size_t nbytes, blocksize, cbytes, nblocks;
  const unsigned char* base;
};

constexpr unsigned kShuffle = 0x1, kMemcpyed = 0x2, kBitShuffle = 0x4, kDontSplit = 0x10;
constexpr size_t kHeader = 16, kMaxSplits = 16, kMinBuffer = 128;

int parse_frame(const unsigned char* src, size_t src_bytes, Frame& f) {
  SRL_REQUIRE(src != nullptr && src_bytes >= kHeader, SRL_ERR_INVALID_ARG,
              "blosc frame: %zu bytes is shorter than the 16-byte header", src_bytes);
  f.flags = src[2];
  f.typesize = src[3];
  f.nbytes = le32(src + 4);
  f.blocksize = le32(src + 8);
  f.cbytes = le32(src + 12);
  f.base = src;
  SRL_REQUIRE(f.cbytes >= kHeader && f.cbytes <= src_bytes, SRL_ERR_INVALID_ARG,
              "blosc frame: header says %zu compressed bytes, the buffer holds %zu", f.cbytes, src_bytes);
  SRL_REQUIRE(f.typesize >= 1, SRL_ERR_INVALID_ARG, "blosc frame: typesize 0");
  f.nblocks = 0;
  if (f.nbytes > 0 && !(f.flags & kMemcpyed)) {
    SRL_REQUIRE(f.blocksize >= 1, SRL_ERR_INVALID_ARG, "blosc frame: blocksize 0");
    f.nblocks = (f.nbytes + f.blocksize - 1) / f.blocksize;
    SRL_REQUIRE(kHeader + 4 * f.nblocks <= f.cbytes, SRL_ERR_INVALID_ARG,
                "blosc frame: %zu block offsets do not fit %zu compressed bytes", f.nblocks, f.cbytes);
  }
  return SRL_OK;
}

// block b of the frame -> dst + b * blocksize; tmp holds one block.  Returns false for malformed input.
bool decode_block(const Frame& f, size_t b, unsigned char* dst, unsigned char* tmp) {
  const size_t lo = b * f.blocksize;
  const size_t bsize = f.nbytes - lo < f.blocksize ? f.nbytes - lo : f.blocksize;
  const bool leftover = bsize != f.blocksize;
  const bool shuffled = (f.flags & kShuffle) && f.typesize > 1;
  const bool split = !(f.flags & kDontSplit) && !leftover && f.typesize <= kMaxSplits && f.blocksize / f.typesize >= kMinBuffer;
  const size_t nsplits = split ? f.typesize : 1;
  const size_t neblock = bsize / nsplits;
  if (neblock * nsplits != bsize) return false;
  const size_t start = le32(f.base + kHeader + 4 * b);
  if (start < kHeader + 4 * f.nblocks || start > f.cbytes) return false;
  const unsigned char* ip = f.base + start;
  const unsigned char* const iend = f.base + f.cbytes;
  unsigned char* out = shuffled ? tmp : dst + lo;
  for (size_t j = 0; j < nsplits; ++j) {
    if (iend - ip < 4) return false;
    const size_t c = le32(ip);
    ip += 4;
    if (c > static_cast<size_t>(iend - ip)) return false;
    if (c == neblock) {
      memcpy(out, ip, neblock);
    } else if (lz4_block_decode(ip, c, out, neblock) != static_cast<long long>(neblock)) {
      return false;
    }
    ip += c;
    out += neblock;
  }
  if (shuffled) {
    const size_t ts = f.typesize, n = bsize / ts;
    unsigned char* d = dst + lo;
    if (ts == 4) {  // the reference's typesize: four byte planes -> elements
      const unsigned char *p0 = tmp, *p1 = tmp + n, *p2 = tmp + 2 * n, *p3 = tmp + 3 * n;
      for (size_t i = 0; i < n; ++i) {
        const uint32_t w = static_cast<uint32_t>(p0[i]) | (static_cast<uint32_t>(p1[i]) << 8) |
                           (static_cast<uint32_t>(p2[i]) << 16) | (static_cast<uint32_t>(p3[i]) << 24);
        memcpy(d + 4 * i, &w, 4);  // little-endian host: byte j of element i
      }
    } else {
      for (size_t j = 0; j < ts; ++j)
        for (size_t i = 0; i < n; ++i) d[i * ts + j] = tmp[j * n + i];
    }
    memcpy(d + n * ts, tmp + n * ts, bsize - n * ts);  // the bytes behind the last whole element
  }
  return true;
}

}  // namespace
}  // namespace srl

extern "C" int srl_lz4_block_decompress(const void* src, size_t src_bytes, void* dst, size_t dst_capacity, size_t* written) {
  using namespace srl;
  SRL_REQUIRE(src != nullptr && src_bytes >= 1 && (dst != nullptr || dst_capacity == 0) && written != nullptr, SRL_ERR_INVALID_ARG,
              "srl_lz4_block_decompress: null pointer or empty input");
  unsigned char none = 0;
  const long long n = lz4_block_decode(static_cast<const unsigned char*>(src), src_bytes,
                                       dst != nullptr ? static_cast<unsigned char*>(dst) : &none, dst_capacity);
  SRL_REQUIRE(n >= 0, SRL_ERR_INVALID_ARG, "srl_lz4_block_decompress: malformed block, or more than %zu bytes of output",
              dst_capacity);
  *written = static_cast<size_t>(n);
  return SRL_OK;
}

extern "C" int srl_blosc1_info(const void* src, size_t src_bytes, size_t* nbytes, size_t* cbytes, size_t* blocksize,
                               int* typesize, int* flags) {
  using namespace srl;
  Frame f;
  const int rc = parse_frame(static_cast<const unsigned char*>(src), src_bytes, f);
  if (rc != SRL_OK) return rc;
  if (nbytes) *nbytes = f.nbytes;
  if (cbytes) *cbytes = f.cbytes;
  if (blocksize) *blocksize = f.blocksize;
  if (typesize) *typesize = static_cast<int>(f.typesize);
  if (flags) *flags = static_cast<int>(f.flags);
  return SRL_OK;
}

extern "C" int srl_blosc1_decompress(const void* src, size_t src_bytes, void* dst, size_t dst_bytes, int threads) {
  using namespace srl;
  Frame f;
  const int rc = parse_frame(static_cast<const unsigned char*>(src), src_bytes, f);
  if (rc != SRL_OK) return rc;
  SRL_REQUIRE(f.nbytes == dst_bytes, SRL_ERR_INVALID_ARG, "srl_blosc1_decompress: the frame holds %zu bytes, the destination %zu",
              f.nbytes, dst_bytes);
  SRL_REQUIRE(threads >= 1 && threads <= 64, SRL_ERR_INVALID_ARG, "srl_blosc1_decompress: threads=%d outside [1, 64]", threads);
  if (f.nbytes == 0) return SRL_OK;
  SRL_REQUIRE(dst != nullptr, SRL_ERR_INVALID_ARG, "srl_blosc1_decompress: null destination");
  unsigned char* out = static_cast<unsigned char*>(dst);
  if (f.flags & kMemcpyed) {
    SRL_REQUIRE(kHeader + f.nbytes <= f.cbytes, SRL_ERR_INVALID_ARG, "blosc frame: stored payload is cut short");
    memcpy(out, f.base + kHeader, f.nbytes);
    return SRL_OK;
  }
  SRL_REQUIRE(!(f.flags & kBitShuffle), SRL_ERR_UNSUPPORTED,
              "blosc frame: bit-shuffled payloads are not decoded here (SRL writes byte-shuffled LZ4, namedarray.py:126)");
  SRL_REQUIRE((f.flags >> 5) == 1, SRL_ERR_UNSUPPORTED,
              "blosc frame: codec family %u; only LZ4 / LZ4HC (1) is decoded here (SRL writes cname='lz4', namedarray.py:126)",
              f.flags >> 5);
  int parts = threads;
  if (static_cast<size_t>(parts) > f.nblocks) parts = static_cast<int>(f.nblocks);
  if (f.nbytes < (1u << 20)) parts = 1;  // a thread start costs more than a small payload
  std::atomic<size_t> next(0);
  std::atomic<bool> ok(true);
  auto work = [&]() {
    std::vector<unsigned char> tmp(f.blocksize < f.nbytes ? f.blocksize : f.nbytes);
    for (size_t b = next.fetch_add(1); b < f.nblocks && ok.load(std::memory_order_relaxed); b = next.fetch_add(1))
      if (!decode_block(f, b, out, tmp.data())) ok.store(false);
  };
  std::vector<std::thread> pool;
  for (int t = 1; t < parts; ++t) pool.emplace_back(work);
  work();
  for (auto& t : pool) t.join();
  SRL_REQUIRE(ok.load(), SRL_ERR_INVALID_ARG, "srl_blosc1_decompress: malformed frame (a block's offsets, sizes or LZ4 stream)");
  return SRL_OK;
}

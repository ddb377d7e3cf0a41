// K5a device code shared by the stand-alone permutation kernel (perm.cu) and the scan kernel that computes the permutation
// on the side (gae_scan_ws.cu).  Spec = oracle/ref_math.py:philox_perm_ref (numpy); bit for bit.
#pragma once
#include "common.cuh"

namespace srl {

struct Philox4 {
  uint32_t c[4];
};

// Philox4x32-10, Salmon et al. SC'11 (Random123); constants as published.
__host__ __device__ __forceinline__ Philox4 philox4x32_10(Philox4 ctr, uint32_t k0, uint32_t k1) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = static_cast<uint64_t>(M0) * ctr.c[0];
    const uint64_t p1 = static_cast<uint64_t>(M1) * ctr.c[2];
    Philox4 n;
    n.c[0] = static_cast<uint32_t>(p1 >> 32) ^ ctr.c[1] ^ k0;
    n.c[1] = static_cast<uint32_t>(p1);
    n.c[2] = static_cast<uint32_t>(p0 >> 32) ^ ctr.c[3] ^ k1;
    n.c[3] = static_cast<uint32_t>(p0);
    ctr = n;
    k0 += W0;
    k1 += W1;
  }
  return ctr;
}

struct PermKeys {
  uint32_t rk[8];
  uint32_t lb, lmask, hmask;
};

__host__ __device__ __forceinline__ uint32_t feistel8(uint32_t x, const PermKeys& k) {
  uint32_t lo = x & k.lmask, hi = (x >> k.lb) & k.hmask;
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    if ((r & 1) == 0) {
      uint32_t f = static_cast<uint32_t>((static_cast<uint64_t>(0xD2511F53u) * (lo ^ k.rk[r])) >> 32);
      f ^= (lo * 0x9E3779B9u) >> 16;
      hi = (hi ^ f) & k.hmask;
    } else {
      uint32_t f = static_cast<uint32_t>((static_cast<uint64_t>(0xCD9E8D57u) * (hi ^ k.rk[r])) >> 32);
      f ^= (hi * 0xBB67AE85u) >> 16;
      lo = (lo ^ f) & k.lmask;
    }
  }
  return (hi << k.lb) | lo;
}


// The eight round keys of epoch `epoch` (two Philox4x32-10 blocks) and the Feistel split of `bits` index bits.
__host__ __device__ __forceinline__ PermKeys perm_keys(uint32_t seed_lo, uint32_t seed_hi, uint32_t epoch, int bits) {
  PermKeys k;
  for (int b = 0; b < 2; ++b) {
    Philox4 ctr;
    ctr.c[0] = b;
    ctr.c[1] = epoch;
    ctr.c[2] = 0x53524C50u;  // 'SRLP'
    ctr.c[3] = 0u;
    const Philox4 o = philox4x32_10(ctr, seed_lo, seed_hi);
    for (int i = 0; i < 4; ++i) k.rk[b * 4 + i] = o.c[i];
  }
  const int lb = bits / 2, hb = bits - lb;
  k.lb = lb;
  k.lmask = (1u << lb) - 1u;
  k.hmask = (1u << hb) - 1u;
  return k;
}

// position e of the permutation of [0, n_env): cycle-walk back into range
__host__ __device__ __forceinline__ uint32_t perm_at(uint32_t e, uint32_t n_env, const PermKeys& k) {
  uint32_t x = feistel8(e, k);
  while (x >= n_env) x = feistel8(x, k);
  return x;
}

// The inverse: the rounds of feistel8 undone in reverse order (each round XORs one half with a function of the other, which
// it leaves alone), and the cycle walk walked backwards.
__host__ __device__ __forceinline__ uint32_t feistel8_inv(uint32_t y, const PermKeys& k) {
  uint32_t lo = y & k.lmask, hi = (y >> k.lb) & k.hmask;
#pragma unroll
  for (int r = 7; r >= 0; --r) {
    if ((r & 1) == 0) {
      uint32_t f = static_cast<uint32_t>((static_cast<uint64_t>(0xD2511F53u) * (lo ^ k.rk[r])) >> 32);
      f ^= (lo * 0x9E3779B9u) >> 16;
      hi = (hi ^ f) & k.hmask;
    } else {
      uint32_t f = static_cast<uint32_t>((static_cast<uint64_t>(0xCD9E8D57u) * (hi ^ k.rk[r])) >> 32);
      f ^= (hi * 0xBB67AE85u) >> 16;
      lo = (lo ^ f) & k.lmask;
    }
  }
  return (hi << k.lb) | lo;
}
// position e with perm_at(e) == x
__host__ __device__ __forceinline__ uint32_t perm_pos_of(uint32_t x, uint32_t n_env, const PermKeys& k) {
  uint32_t e = feistel8_inv(x, k);
  while (e >= n_env) e = feistel8_inv(e, k);
  return e;
}

inline int perm_bits(int n_env) {
  int bits = 0;
  while ((1ll << bits) < n_env) ++bits;  // == (n_env - 1).bit_length()
  return bits < 2 ? 2 : bits;
}

// A permutation to be computed on the side by another kernel (gae_scan_ws.cu): out == nullptr -> none.
struct PermJob {
  uint32_t seed_lo, seed_hi, epoch0;
  int n_epochs, n_env, group, bits;
  int32_t* out;
  // optional (part != nullptr): every scan CTA also adds ITS lanes' {count, sum, sum of squares} per minibatch -- epoch e's
  // minibatch j = positions [j * per_mb, (j + 1) * per_mb) of the permuted lane list -- into part[e * minibatches + j][cta][4]
  int minibatches, per_mb;
  double* part;
};
constexpr int kPartSlots = 32;   // minibatches of a step the partial sums cover (== SRL_MAX_LOSS_BATCH)
constexpr int kPartEpochs = 8;

}  // namespace srl

"""TEST DOUBLE: a writer of Blosc-1 frames (what `blosc.compress(payload, typesize=4, cname='lz4')` produces), following the
published layout -- c-blosc README_HEADER.rst and blosc.c -- with liblz4 itself (pyarrow's `lz4_raw` codec) compressing the
streams.  `blosc` is not installed in this image, so the native decoder (srl_b200/csrc/blosc_decode.cu) cannot be checked against
frames the real package wrote: its LZ4 layer is pinned to liblz4 through this file, its framing is checked against this
restatement of the same description (UNPINNED against blosc; srl_b200/wire.py cross-checks the decoder against the real package
at run time wherever it is importable).

  header  : version 2, versionlz 1, flags, typesize, nbytes u32, blocksize u32, cbytes u32 (little endian)
  flags   : 0x01 byte shuffle, 0x02 memcpyed (payload stored), 0x10 blocks not split, (codec family << 5), LZ4 = 1
  bstarts : one int32 per block, offset of the block's streams from the start of the frame
  block   : nsplits x {int32 compressed size, bytes}; compressed size == decoded size means the stream is stored
  split   : typesize streams per block unless 0x10, the shorter last block, typesize > 16 or blocksize / typesize < 128
  shuffle : within a block of n whole elements, byte j of element i moves to j * n + i; trailing bytes stay
"""
import struct

import numpy as np
import pyarrow as pa

SHUFFLE, MEMCPYED, DONT_SPLIT, LZ4_FAMILY = 0x01, 0x02, 0x10, 1 << 5
_lz4 = pa.Codec("lz4_raw")


def lz4_block(data: bytes) -> bytes:
    """One raw LZ4 block written by liblz4."""
    return _lz4.compress(bytes(data), asbytes=True)


def shuffle_block(block: bytes, typesize: int) -> bytes:
    n = len(block) // typesize
    body = np.frombuffer(block, dtype=np.uint8, count=n * typesize).reshape(n, typesize).T.tobytes()
    return body + block[n * typesize:]


def compress(payload: bytes, typesize: int = 4, blocksize: int = 0, shuffle: bool = True, split: bool = True,
             memcpyed: bool = False, store_incompressible: bool = True) -> bytes:
    payload = bytes(payload)
    nbytes = len(payload)
    blocksize = max(1, min(blocksize or (1 << 16), max(nbytes, 1)))
    if blocksize > typesize:  # blosc keeps a block a whole number of elements (compute_blocksize)
        blocksize = blocksize // typesize * typesize
    flags = LZ4_FAMILY | (SHUFFLE if shuffle else 0) | (0 if split else DONT_SPLIT)
    if memcpyed:
        return struct.pack("<BBBBIII", 2, 1, flags | MEMCPYED, typesize, nbytes, blocksize, 16 + nbytes) + payload
    nblocks = -(-nbytes // blocksize) if nbytes else 0
    body, bstarts = b"", []
    for b in range(nblocks):
        block = payload[b * blocksize:(b + 1) * blocksize]
        leftover = len(block) != blocksize
        if shuffle and typesize > 1:
            block = shuffle_block(block, typesize)
        nsplits = typesize if (split and not leftover and typesize <= 16 and blocksize // typesize >= 128) else 1
        ne = len(block) // nsplits
        bstarts.append(16 + 4 * nblocks + len(body))
        for j in range(nsplits):
            stream = block[j * ne:(j + 1) * ne]
            c = lz4_block(stream)
            if store_incompressible and len(c) >= ne:
                c = stream  # compressed size == decoded size: stored
            assert len(c) != ne or c == stream
            body += struct.pack("<i", len(c)) + c
    head = struct.pack("<BBBBIII", 2, 1, flags, typesize, nbytes, blocksize, 16 + 4 * nblocks + len(body))
    return head + b"".join(struct.pack("<i", s) for s in bstarts) + body

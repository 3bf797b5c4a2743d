// zk_sha256.cuh — SHA-256 for the Fiat-Shamir challenges.
//
// Replaces WebCrypto `crypto.subtle.digest('SHA-256', bytes)` inside hashPoints
// (/root/reference/src/curves/group.ts:221-233).  The message is the concatenation of
// 65/67-byte point encodings gathered from several places, so the hasher is streaming:
// update() with arbitrary byte spans, final80() returns the first 10 digest bytes as the
// 80-bit challenge (two words: hi 16 bits, lo 64 bits).
#pragma once
#include "zk_field.cuh"

namespace zk {

struct Sha256 {
  uint32_t h[8];
  uint32_t w[16];   // current block, big-endian words
  uint32_t fill;    // bytes in current block
  uint64_t total;   // total bytes

  ZK_HD static uint32_t rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
  ZK_HD static uint32_t K(int i) {
    constexpr uint32_t k[64] = {
        0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5,
        0xd807aa98, 0x12835b01, 0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174,
        0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da,
        0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967,
        0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
        0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070,
        0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3,
        0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};
    return k[i];
  }

  ZK_HD void init() {
    h[0] = 0x6a09e667; h[1] = 0xbb67ae85; h[2] = 0x3c6ef372; h[3] = 0xa54ff53a;
    h[4] = 0x510e527f; h[5] = 0x9b05688c; h[6] = 0x1f83d9ab; h[7] = 0x5be0cd19;
#pragma unroll
    for (int i = 0; i < 16; i++) w[i] = 0;
    fill = 0;
    total = 0;
    cur = 0;
  }

  ZK_HD void compress() {
    uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
    uint32_t m[16];
#pragma unroll
    for (int i = 0; i < 16; i++) m[i] = w[i];
#pragma unroll
    for (int i = 0; i < 64; i++) {
      uint32_t wi;
      if (i < 16) {
        wi = m[i];
      } else {
        uint32_t w15 = m[(i + 1) & 15], w2 = m[(i + 14) & 15];
        uint32_t s0 = rotr(w15, 7) ^ rotr(w15, 18) ^ (w15 >> 3);
        uint32_t s1 = rotr(w2, 17) ^ rotr(w2, 19) ^ (w2 >> 10);
        wi = m[i & 15] + s0 + m[(i + 9) & 15] + s1;
        m[i & 15] = wi;
      }
      uint32_t S1 = rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25);
      uint32_t ch = (e & f) ^ (~e & g);
      uint32_t t1 = hh + S1 + ch + K(i) + wi;
      uint32_t S0 = rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22);
      uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
      uint32_t t2 = S0 + mj;
      hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
    }
    h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
    fill -= 64;   // bytes still pending in `cur` (0..3) belong to the next block
  }

  // The block buffer is a 16-word FIFO with static indices only (w[15] is the newest word), so it
  // stays in registers; `fill` counts bytes of the current block, `cur` gathers 4 bytes.
  uint32_t cur;
  ZK_HD void push_word(uint32_t v) {
#pragma unroll
    for (int i = 0; i < 15; i++) w[i] = w[i + 1];
    w[15] = v;
  }
  ZK_HD void put(uint8_t byte) {
    cur = (cur << 8) | byte;
    fill++;
    total++;
    if ((fill & 3) == 0) {
      push_word(cur);
      if (fill == 64) compress();
    }
  }
  // four stream bytes at once, given as the little-endian word loaded from memory
  ZK_HD void put4(uint32_t lw) {
    const uint32_t bw = (lw << 24) | ((lw & 0xff00u) << 8) | ((lw >> 8) & 0xff00u) | (lw >> 24);
    const int nb = (int)(fill & 3);                       // bytes pending in `cur`
    const uint64_t v = ((uint64_t)cur << 32) | bw;
    push_word((uint32_t)(v >> (8 * nb)));
    cur = bw;                                             // its low nb bytes are the new pending bytes
    fill += 4;
    total += 4;
    if (fill >= 64) compress();
  }
  ZK_HD void update(const uint8_t* p, int n) {
    int i = 0;
    while (i < n && (((size_t)(p + i)) & 3)) put(p[i++]);
    for (; i + 4 <= n; i += 4) put4(*reinterpret_cast<const uint32_t*>(p + i));
    for (; i < n; i++) put(p[i]);
  }
  // digest[0..9] as (hi16, lo64): challenge = hi16 * 2^64 + lo64
  ZK_HD void final80(uint32_t* c3) {  // c3[0] = low 32, c3[1] = mid 32, c3[2] = top 16 bits
    const uint64_t bits = total * 8;
    put(0x80);
    while ((fill & 3) != 0) put(0);
    // now `fill` is a multiple of 4; pad with zero words up to byte 56, then the length
    if (fill > 56) {
      while (fill != 0) { push_word(0); fill += 4; if (fill == 64) compress(); }
    }
    while (fill != 56) { push_word(0); fill += 4; }
    push_word((uint32_t)(bits >> 32));
    push_word((uint32_t)bits);
    compress();
    // digest bytes 0..9 = h0 (4) h1 (4) top half of h2 (2)  -> 80-bit big-endian integer
    uint32_t top16 = h[0] >> 16;
    uint32_t mid = (h[0] << 16) | (h[1] >> 16);
    uint32_t low = (h[1] << 16) | (h[2] >> 16);
    c3[0] = low;
    c3[1] = mid;
    c3[2] = top16;
  }
};

}  // namespace zk

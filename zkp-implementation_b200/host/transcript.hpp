// Fiat-Shamir transcript of the reference's PLONK: plonk/src/challenge.rs:49-89.
//
//   feed(c):  data = SHA-256(data_or_empty || serialize_uncompressed(c))
//   generate_challenges<N>():  seed = u64::from_le_bytes(data[0..8]); StdRng::seed_from_u64(seed);
//                              N x Fr::rand(&mut rng); panics if called twice without a feed
//
// The pieces below restate the third-party semantics the reference relies on (SURVEY.md App. A.5/A.6;
// they cannot be cross-checked against a Rust run in this image):
//   * ark-bls12-381 0.4 G1 `serialize_uncompressed`: x || y as 48-byte big-endian canonical integers,
//     infinity = 0x40 followed by 95 zero bytes
//   * rand_core 0.6 `seed_from_u64`: PCG32 expansion of the u64 into the 32-byte ChaCha key
//   * rand 0.8 `StdRng` = ChaCha12, 64-bit block counter from 0, stream 0, u64 = two consecutive words
//   * ark-ff 0.4 `Fr::rand`: 4 x next_u64 are the limbs of the Montgomery representation, top bit
//     cleared, redraw while >= r
#pragma once
#include <stdint.h>
#include <string.h>

#include <vector>

#include "mont_host.hpp"

namespace zkp_host {

// ---- SHA-256 (FIPS 180-4) ----------------------------------------------------------------------
class Sha256 {
 public:
  Sha256() { reset(); }
  void reset() {
    static const uint32_t iv[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
    memcpy(h_, iv, sizeof(iv));
    len_ = 0;
    fill_ = 0;
  }
  void update(const uint8_t* d, size_t n) {
    len_ += n;
    while (n) {
      size_t take = 64 - fill_ < n ? 64 - fill_ : n;
      memcpy(buf_ + fill_, d, take);
      fill_ += take; d += take; n -= take;
      if (fill_ == 64) { block(buf_); fill_ = 0; }
    }
  }
  void finalize(uint8_t out[32]) {
    uint64_t bits = len_ * 8;
    uint8_t pad = 0x80;
    update(&pad, 1);
    uint8_t z = 0;
    while (fill_ != 56) update(&z, 1);
    uint8_t lenb[8];
    for (int i = 0; i < 8; i++) lenb[i] = (uint8_t)(bits >> (56 - 8 * i));
    update(lenb, 8);
    for (int i = 0; i < 8; i++) { out[4 * i] = h_[i] >> 24; out[4 * i + 1] = h_[i] >> 16; out[4 * i + 2] = h_[i] >> 8; out[4 * i + 3] = h_[i]; }
  }

 private:
  static uint32_t rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
  void block(const uint8_t* p) {
    static const uint32_t K[64] = {
        0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01,
        0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc,
        0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147,
        0x06ca6351, 0x14292967, 0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
        0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08,
        0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208,
        0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};
    uint32_t w[64];
    for (int i = 0; i < 16; i++) w[i] = (uint32_t)p[4 * i] << 24 | (uint32_t)p[4 * i + 1] << 16 | (uint32_t)p[4 * i + 2] << 8 | p[4 * i + 3];
    for (int i = 16; i < 64; i++) {
      uint32_t s0 = rotr(w[i - 15], 7) ^ rotr(w[i - 15], 18) ^ (w[i - 15] >> 3);
      uint32_t s1 = rotr(w[i - 2], 17) ^ rotr(w[i - 2], 19) ^ (w[i - 2] >> 10);
      w[i] = w[i - 16] + s0 + w[i - 7] + s1;
    }
    uint32_t a = h_[0], b = h_[1], c = h_[2], d = h_[3], e = h_[4], f = h_[5], g = h_[6], h = h_[7];
    for (int i = 0; i < 64; i++) {
      uint32_t S1 = rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25);
      uint32_t ch = (e & f) ^ (~e & g);
      uint32_t t1 = h + S1 + ch + K[i] + w[i];
      uint32_t S0 = rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22);
      uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
      uint32_t t2 = S0 + mj;
      h = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
    }
    h_[0] += a; h_[1] += b; h_[2] += c; h_[3] += d; h_[4] += e; h_[5] += f; h_[6] += g; h_[7] += h;
  }
  uint32_t h_[8];
  uint8_t buf_[64];
  uint64_t len_;
  size_t fill_;
};

// ---- StdRng::seed_from_u64 + ChaCha12 ------------------------------------------------------------
class StdRng {
 public:
  // PCG32 (XSH-RR) output function of rand_core's seed expansion; also the output function of the public pcg32
  // generator, which is what tests/test_transcript_kat.py pins it against.
  static uint32_t pcg32_output(uint64_t state) {
    uint32_t xorshifted = (uint32_t)(((state >> 18) ^ state) >> 27);
    uint32_t rot = (uint32_t)(state >> 59);
    return (xorshifted >> rot) | (xorshifted << ((32 - rot) & 31));
  }
  // rand_core 0.6 `SeedableRng::seed_from_u64`: eight PCG32 steps (advance, then output) -> 32-byte seed
  static void seed_from_u64(uint64_t state, uint32_t key[8]) {
    for (int i = 0; i < 8; i++) {
      state = state * 6364136223846793005ull + 11634580027462260723ull;
      key[i] = pcg32_output(state);
    }
  }
  explicit StdRng(uint64_t state) {
    seed_from_u64(state, key_);
    ctr_ = 0;
    pos_ = 16;
  }
  // `StdRng::from_seed`: the 32 seed bytes are the eight little-endian key words.  double_rounds = 6 is ChaCha12
  // (rand 0.8's StdRng); 10 gives ChaCha20 so the block function can be checked against RFC 8439 vectors.
  explicit StdRng(const uint32_t key[8], int double_rounds = 6) : double_rounds_(double_rounds) {
    for (int i = 0; i < 8; i++) key_[i] = key[i];
    ctr_ = 0;
    pos_ = 16;
  }
  uint32_t next_u32() {
    if (pos_ == 16) { block(); pos_ = 0; }
    return out_[pos_++];
  }
  uint64_t next_u64() {
    uint64_t lo = next_u32();
    uint64_t hi = next_u32();
    return lo | (hi << 32);
  }
  Fr fr_rand() {
    for (;;) {
      Fr x;
      for (int i = 0; i < 4; i++) x.v[i] = next_u64();
      x.v[3] &= 0x7fffffffffffffffull;   // 256 - 255 bits shaved
      if (!geq<4>(x.v, FR().p)) return x;  // limbs are taken as the Montgomery representation
    }
  }

 private:
  static uint32_t rotl(uint32_t x, int n) { return (x << n) | (x >> (32 - n)); }
  void block() {
    uint32_t st[16] = {0x61707865, 0x3320646e, 0x79622d32, 0x6b206574};
    for (int i = 0; i < 8; i++) st[4 + i] = key_[i];
    st[12] = (uint32_t)ctr_; st[13] = (uint32_t)(ctr_ >> 32); st[14] = 0; st[15] = 0;
    uint32_t x[16];
    memcpy(x, st, sizeof(x));
#define ZKP_QR(a, b, c, d) \
  x[a] += x[b]; x[d] = rotl(x[d] ^ x[a], 16); x[c] += x[d]; x[b] = rotl(x[b] ^ x[c], 12); \
  x[a] += x[b]; x[d] = rotl(x[d] ^ x[a], 8);  x[c] += x[d]; x[b] = rotl(x[b] ^ x[c], 7);
    for (int r = 0; r < double_rounds_; r++) {
      ZKP_QR(0, 4, 8, 12) ZKP_QR(1, 5, 9, 13) ZKP_QR(2, 6, 10, 14) ZKP_QR(3, 7, 11, 15)
      ZKP_QR(0, 5, 10, 15) ZKP_QR(1, 6, 11, 12) ZKP_QR(2, 7, 8, 13) ZKP_QR(3, 4, 9, 14)
    }
#undef ZKP_QR
    for (int i = 0; i < 16; i++) out_[i] = x[i] + st[i];
    ctr_++;
  }
  uint32_t key_[8];
  uint32_t out_[16];
  uint64_t ctr_;
  int pos_;
  int double_rounds_ = 6;
};

// ---- G1 point as it crosses the C ABI: x || y Montgomery limbs, (0,0) = infinity ------------------
struct G1 {
  uint64_t xy[12];
  bool is_inf() const {
    uint64_t o = 0;
    for (int i = 0; i < 12; i++) o |= xy[i];
    return o == 0;
  }
};

inline void g1_serialize_uncompressed(const G1& p, uint8_t out[96]) {
  memset(out, 0, 96);
  if (p.is_inf()) { out[0] = 0x40; return; }
  const uint64_t one[6] = {1, 0, 0, 0, 0, 0};
  for (int c = 0; c < 2; c++) {
    uint64_t canon[6];
    mont_mul<6>(FQ(), canon, p.xy + 6 * c, one);  // out of Montgomery form
    for (int i = 0; i < 48; i++) out[48 * c + i] = (uint8_t)(canon[5 - i / 8] >> (56 - 8 * (i % 8)));
  }
}

class ChallengeGenerator {
 public:
  void feed(const G1& c) {
    Sha256 h;
    if (have_) h.update(data_, 32);
    uint8_t ser[96];
    g1_serialize_uncompressed(c, ser);
    h.update(ser, 96);
    h.finalize(data_);
    have_ = true;
    generated_ = false;
  }
  // returns false where the reference panics ("I'm hungry! Feed me something first" / no data)
  bool generate(int n, Fr* out) {
    if (generated_ || !have_) return false;
    generated_ = true;
    uint64_t seed = 0;
    for (int i = 0; i < 8; i++) seed |= (uint64_t)data_[i] << (8 * i);
    StdRng rng(seed);
    for (int i = 0; i < n; i++) out[i] = rng.fr_rand();
    return true;
  }

 private:
  uint8_t data_[32];
  bool have_ = false;
  bool generated_ = false;
};

}  // namespace zkp_host

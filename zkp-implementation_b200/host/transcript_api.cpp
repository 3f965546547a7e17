// C entry points over host/transcript.hpp for the known-answer tests (include/zkp_plonk.h, last section).
#include "../../include/zkp_plonk.h"
#include "transcript.hpp"

using namespace zkp_host;

extern "C" {

void zkp_transcript_sha256(const uint8_t* data, size_t n, uint8_t out[32]) {
  Sha256 h;
  h.update(data, n);
  h.finalize(out);
}

uint32_t zkp_transcript_pcg32_output(uint64_t state) { return StdRng::pcg32_output(state); }

void zkp_transcript_seed_from_u64(uint64_t seed, uint32_t key_out[8]) { StdRng::seed_from_u64(seed, key_out); }

void zkp_transcript_chacha_words(const uint32_t key[8], int double_rounds, size_t count, uint32_t* out) {
  StdRng rng(key, double_rounds);
  for (size_t i = 0; i < count; i++) out[i] = rng.next_u32();
}

void zkp_transcript_g1_serialize(const uint64_t xy[12], uint8_t out[96]) {
  G1 p;
  memcpy(p.xy, xy, sizeof(p.xy));
  g1_serialize_uncompressed(p, out);
}

int zkp_transcript_challenges(const uint64_t* points, size_t k, size_t n, uint64_t* out) {
  ChallengeGenerator ch;
  for (size_t i = 0; i < k; i++) {
    G1 p;
    memcpy(p.xy, points + 12 * i, sizeof(p.xy));
    ch.feed(p);
  }
  std::vector<Fr> c(n);
  if (!ch.generate((int)n, c.data())) return ZKP_PLONK_ERR_TRANSCRIPT;
  for (size_t i = 0; i < n; i++) memcpy(out + 4 * i, c[i].v, 32);
  return 0;
}

}  // extern "C"

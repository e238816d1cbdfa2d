/* ORACLE (test infrastructure, NOT product code).
 *
 * Bit-exact CPU restatement of the VQ codebook lookup of the reference path:
 * diffusers VectorQuantizer.forward -> argmin(cdist(z, E)) called at
 * ivideogpt/vq_model/compressive_vq_model.py:199,202.
 *
 * The arithmetic ORDER is part of the definition shared with ivideogpt_b200/csrc/vq_argmin.cu so that the int64
 * result can be compared for exact equality:
 *     dot(n,k)   = fmaf chain over d = 0..D-1, starting at +0.0f
 *     enorm(k)   = fmaf chain over d of e*e
 *     score(n,k) = fmaf(-2, dot, enorm)            (= ||z-e||^2 - ||z||^2; per-row monotone map of cdist)
 *     idx(n)     = lowest k among the minima       (torch.argmin tie rule)
 * tests/test_vq_argmin.py additionally checks agreement with torch.cdist+argmin (the literal reference
 * expression) on every row whose top-2 margin is above float rounding noise.
 * Build: gcc -O2 -ffp-contract=off -shared -fPIC (see oracle/Makefile); fmaf is the correctly rounded libm/FMA.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>

void vq_argmin_ref(const float* z, const float* e, int64_t* idx, float* best_out, float* second_out, int N, int K,
                   int D) {
  for (int n = 0; n < N; ++n) {
    float best = INFINITY, second = INFINITY;
    int64_t bi = 0;
    for (int k = 0; k < K; ++k) {
      float dot = 0.0f, en = 0.0f;
      for (int d = 0; d < D; ++d) {
        dot = fmaf(z[(size_t)n * D + d], e[(size_t)k * D + d], dot);
        en = fmaf(e[(size_t)k * D + d], e[(size_t)k * D + d], en);
      }
      float s = fmaf(-2.0f, dot, en);
      if (s < best) { second = best; best = s; bi = k; }
      else if (s < second) { second = s; }
    }
    idx[n] = bi;
    if (best_out) best_out[n] = best;
    if (second_out) second_out[n] = second;
  }
}

/* order 1: even-d and odd-d chains (the packed fma.rn.f32x2 kernel):  dot = chain_even + chain_odd */
void vq_argmin_ref_order1(const float* z, const float* e, int64_t* idx, float* best_out, float* second_out, int N,
                          int K, int D) {
  for (int n = 0; n < N; ++n) {
    float best = INFINITY, second = INFINITY;
    int64_t bi = 0;
    for (int k = 0; k < K; ++k) {
      float ev = 0.0f, od = 0.0f, en = 0.0f;
      for (int d = 0; d < D; d += 2) {
        ev = fmaf(z[(size_t)n * D + d], e[(size_t)k * D + d], ev);
        od = fmaf(z[(size_t)n * D + d + 1], e[(size_t)k * D + d + 1], od);
      }
      for (int d = 0; d < D; ++d) en = fmaf(e[(size_t)k * D + d], e[(size_t)k * D + d], en);
      float s = fmaf(-2.0f, ev + od, en);
      if (s < best) { second = best; best = s; bi = k; }
      else if (s < second) { second = s; }
    }
    idx[n] = bi;
    if (best_out) best_out[n] = best;
    if (second_out) second_out[n] = second;
  }
}

/* psum_oracle.c -- plain-C CPU restatement of the Pauli-sum matvec.  TEST INFRASTRUCTURE ONLY.
 *
 * Used by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs as
 * the checker and the timed CPU baseline; never by the product path (qiskit_aqua_b200).
 * Parity status: PINNED -- tests/test_oracle_golden.py checks this file against
 * oracle/pauli_oracle.py, which is itself pinned by the reference's golden vectors G1..G13.
 *
 * Arithmetic restated: the reference sums one CSR per Pauli
 * (qiskit/aqua/operators/legacy/op_converter.py:120-123) where terra's Pauli.to_matrix gives
 *     P[i, i^x] = (-i)^{popc(x&z)} (-1)^{popc(i&z)},
 * then evaluates np.vdot(psi, H.dot(psi)) (legacy/weighted_pauli_operator.py:621-622).
 * Row i of H.psi is therefore  y[i] = sum_k c_k (-i)^{ypow_k} (-1)^{popc(i & z_k)} psi[i ^ x_k].
 */
#include <stdint.h>
#include <stdlib.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static inline int parity64(uint64_t v) { return __builtin_popcountll(v) & 1; }

static inline uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
/* twin of counter_uniform() in qiskit_aqua_b200/csrc/psum_kernels.cuh */
static inline double counter_uniform(uint64_t seed, uint64_t ctr) {
  uint64_t h = splitmix64(seed ^ splitmix64(ctr));
  return (double)(h >> 11) * (1.0 / 9007199254740992.0) - 0.5;
}

/* torchrun exports OMP_NUM_THREADS=1; the CPU baseline must still use every host core */
void psum_oracle_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

int psum_oracle_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* out[2*j], out[2*j+1] = amplitude first+j of the synthetic state */
void psum_oracle_state(uint64_t seed, uint64_t first, uint64_t count, double* out) {
#pragma omp parallel for schedule(static)
  for (int64_t j = 0; j < (int64_t)count; ++j) {
    uint64_t gi = first + (uint64_t)j;
    out[2 * j] = counter_uniform(seed, 2 * gi);
    out[2 * j + 1] = counter_uniform(seed, 2 * gi + 1);
  }
}

static inline void phase_coeff(const double* c, const uint8_t* ypow, int64_t k, double* re, double* im) {
  double a = c[2 * k], b = c[2 * k + 1];
  switch (ypow ? (ypow[k] & 3) : 0) {
    case 1: *re = b; *im = -a; break;   /* * (-i)  */
    case 2: *re = -a; *im = -b; break;  /* * (-1)  */
    case 3: *re = -b; *im = a; break;   /* * (+i)  */
    default: *re = a; *im = b; break;
  }
}

/* y[i - row_begin] for rows [row_begin, row_end); psi has 2^n amplitudes (interleaved). */
void psum_oracle_apply(int n, int64_t T, const uint64_t* x, const uint64_t* z, const uint8_t* ypow,
                       const double* coeff, const double* psi, double* y, uint64_t row_begin,
                       uint64_t row_end) {
  (void)n;
  double* cre = (double*)malloc(sizeof(double) * (size_t)(T > 0 ? T : 1));
  double* cim = (double*)malloc(sizeof(double) * (size_t)(T > 0 ? T : 1));
  for (int64_t k = 0; k < T; ++k) phase_coeff(coeff, ypow, k, &cre[k], &cim[k]);
  const uint64_t B = 4096; /* rows per block: the block's y stays in L2 while terms stream by */
  if (row_begin % B == 0 && row_end % B == 0 && row_end > row_begin) {
    /* term-outer, row-block-inner order: for a fixed term the partner amplitudes i^x of an
       aligned block form one contiguous aligned block, so memory is streamed, not gathered.
       Per row the sum over k runs in the same order as the row-major loop below. */
    const int64_t nblk = (int64_t)((row_end - row_begin) / B);
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t b = 0; b < nblk; ++b) {
      const uint64_t base = row_begin + (uint64_t)b * B;
      double* yb = y + 2 * ((uint64_t)b * B);
      for (uint64_t t = 0; t < 2 * B; ++t) yb[t] = 0.0;
      for (int64_t k = 0; k < T; ++k) {
        const uint64_t xk = x[k], zk = z[k];
        const double cr = cre[k], ci = cim[k];
        const double* pb = psi + 2 * ((base ^ xk) & ~(B - 1));
        const uint64_t xl = xk & (B - 1);
        const int pbase = parity64(base & zk);
        const uint64_t zl = zk & (B - 1);
        for (uint64_t t = 0; t < B; ++t) {
          const uint64_t j = t ^ xl;
          const double pr = pb[2 * j], pi = pb[2 * j + 1];
          double tr = cr * pr - ci * pi;
          double ti = cr * pi + ci * pr;
          if (pbase ^ parity64(t & zl)) { tr = -tr; ti = -ti; }
          yb[2 * t] += tr;
          yb[2 * t + 1] += ti;
        }
      }
    }
    free(cre);
    free(cim);
    return;
  }
#pragma omp parallel for schedule(static)
  for (int64_t r = (int64_t)row_begin; r < (int64_t)row_end; ++r) {
    const uint64_t i = (uint64_t)r;
    double are = 0.0, aim = 0.0;
    for (int64_t k = 0; k < T; ++k) {
      const uint64_t j = i ^ x[k];
      const double pr = psi[2 * j], pi = psi[2 * j + 1];
      double tr = cre[k] * pr - cim[k] * pi;
      double ti = cre[k] * pi + cim[k] * pr;
      if (parity64(i & z[k])) { tr = -tr; ti = -ti; }
      are += tr;
      aim += ti;
    }
    y[2 * (i - row_begin)] = are;
    y[2 * (i - row_begin) + 1] = aim;
  }
  free(cre);
  free(cim);
}

/* same, at an arbitrary list of rows, with psi = the counter-based synthetic state scaled by
 * `scale` (no 2^n host array needed: used for sampled parity checks at 30+ qubits). */
void psum_oracle_apply_rows_counter(int64_t T, const uint64_t* x, const uint64_t* z, const uint8_t* ypow,
                                    const double* coeff, uint64_t seed, double scale,
                                    const uint64_t* rows, int64_t nrows, double* y) {
  double* cre = (double*)malloc(sizeof(double) * (size_t)(T > 0 ? T : 1));
  double* cim = (double*)malloc(sizeof(double) * (size_t)(T > 0 ? T : 1));
  for (int64_t k = 0; k < T; ++k) phase_coeff(coeff, ypow, k, &cre[k], &cim[k]);
#pragma omp parallel for schedule(dynamic, 4)
  for (int64_t r = 0; r < nrows; ++r) {
    const uint64_t i = rows[r];
    double are = 0.0, aim = 0.0;
    for (int64_t k = 0; k < T; ++k) {
      const uint64_t j = i ^ x[k];
      const double pr = scale * counter_uniform(seed, 2 * j), pi = scale * counter_uniform(seed, 2 * j + 1);
      double tr = cre[k] * pr - cim[k] * pi;
      double ti = cre[k] * pi + cim[k] * pr;
      if (parity64(i & z[k])) { tr = -tr; ti = -ti; }
      are += tr;
      aim += ti;
    }
    y[2 * r] = are;
    y[2 * r + 1] = aim;
  }
  free(cre);
  free(cim);
}

/* <psi|H|psi> = sum_i conj(psi_i) y_i over all 2^n rows */
void psum_oracle_expect(int n, int64_t T, const uint64_t* x, const uint64_t* z, const uint8_t* ypow,
                        const double* coeff, const double* psi, double* out2) {
  const uint64_t N = 1ull << n;
  double* cre = (double*)malloc(sizeof(double) * (size_t)(T > 0 ? T : 1));
  double* cim = (double*)malloc(sizeof(double) * (size_t)(T > 0 ? T : 1));
  for (int64_t k = 0; k < T; ++k) phase_coeff(coeff, ypow, k, &cre[k], &cim[k]);
  double sre = 0.0, sim = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : sre, sim)
  for (int64_t r = 0; r < (int64_t)N; ++r) {
    const uint64_t i = (uint64_t)r;
    double are = 0.0, aim = 0.0;
    for (int64_t k = 0; k < T; ++k) {
      const uint64_t j = i ^ x[k];
      const double pr = psi[2 * j], pi = psi[2 * j + 1];
      double tr = cre[k] * pr - cim[k] * pi;
      double ti = cre[k] * pi + cim[k] * pr;
      if (parity64(i & z[k])) { tr = -tr; ti = -ti; }
      are += tr;
      aim += ti;
    }
    sre += psi[2 * i] * are + psi[2 * i + 1] * aim;
    sim += psi[2 * i] * aim - psi[2 * i + 1] * are;
  }
  out2[0] = sre;
  out2[1] = sim;
  free(cre);
  free(cim);
}

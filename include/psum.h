/* psum.h -- C ABI of libpsum.so: the B200-native Pauli-sum H.psi / <psi|H|psi> engine.
 *
 * The reference (qiskit-aqua 0.10.0) is pure Python and has NO native interface for this path
 * (its setup.py:38-76 declares no extension modules), so there is no existing FFI to match.
 * Each entry point below names the reference function whose arithmetic it replaces; the Python
 * shim in qiskit_aqua_b200/ binds them with ctypes (INTEGRATION.md shows the stub a maintainer
 * would add on the reference side).
 *
 * Conventions
 *   - n qubits, N = 2^n amplitudes, qubit q <-> bit q of the basis index.
 *   - a term k is (x_mask, z_mask, ypow, coeff):  P_k[i, i^x] = (-i)^ypow (-1)^popc(i & z)
 *     with ypow = popc(x & z) mod 4 for a label Pauli (terra Pauli.to_matrix convention).
 *   - state vectors are complex128 (interleaved re,im doubles), DEVICE pointers, 16-byte aligned.
 *   - every function returns PSUM_OK (0) or a negative code; psum_last_error() gives the text
 *     (thread-local).  No exceptions/aborts cross the ABI.  There is NO CPU fallback: compute
 *     entry points fail with PSUM_ERR_CUDA when no device is usable.
 *   - a handle is not thread-safe; calls are ordered on the given stream; calls that return a
 *     scalar to the host synchronise that stream.
 */
#ifndef PSUM_H
#define PSUM_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct psum_handle_s* psum_handle_t;
typedef void* psum_stream_t; /* cudaStream_t */

#define PSUM_OK 0
#define PSUM_ERR_INVALID (-1)
#define PSUM_ERR_CUDA (-2)
#define PSUM_ERR_EMPTY (-3)          /* reference: AquaError("Operator is empty") weighted_pauli_operator.py:616 */
#define PSUM_ERR_NOT_HERMITIAN (-4)  /* Lanczos needs real label coefficients */
#define PSUM_ERR_NOT_DIAGONAL (-5)
#define PSUM_ERR_NCCL (-6)
#define PSUM_ERR_NOMEM (-7)
#define PSUM_ERR_NOT_CONVERGED (-8)

const char* psum_last_error(void);
int psum_version(void);

/* ---- operator handle ------------------------------------------------------------------ */
/* Pack step.  Replaces WeightedPauliOperator.__init__/simplify
 * (legacy/weighted_pauli_operator.py:43-70, 341-400): duplicate (x,z) are merged, exact zeros
 * dropped; and op_converter.to_matrix_operator (legacy/op_converter.py:103-124) in the sense that
 * no matrix is ever built.  Host-only work; the device copy of the plan is made on first use. */
int psum_create(int n_qubits, int64_t n_terms, const uint64_t* x_mask, const uint64_t* z_mask,
                const uint8_t* ypow, const double* coeff_re_im, psum_handle_t* out);
int psum_destroy(psum_handle_t h);

/* Tuning/testing knobs (before first use): "tile_bits" (11..13), "low_bits" (1..8),
 * "kernel" (0 auto, 1 streaming kernel A only, 2 tiled kernel B), "min_cover" (3: a pass is opened only
 * for that many x-groups, fewer read their partners from global memory), "plan_tries" (64: randomised
 * restarts of the pass planner, the plan with the least HBM traffic wins);
 * "reg_bits" (3 default | 4): amplitudes per thread = 2^reg_bits.  x-groups whose masks differ only
 * in the register qubits of a pass share one set of partner loads (a "bundle");
 * "cta_mode" (0 auto | 1 two CTAs per SM | 2 one big CTA per SM): a pass with quadratic groups or more
 * than one chunk runs as one CTA per SM with twice the threads and large staging capacities;
 * "term_cache" (1 default): one-chunk passes without imaginary patterns keep their component terms in shared
 * memory when they fit, so the per-tile fold does not read them from L2;
 * "copy_threads" (0 auto): host threads that fill the pinned bounce ring for pageable input;
 * "herm_pairing" (1 default): psum_expect of a Hermitian operator evaluates each amplitude pair
 * (i, i^x) once and doubles it (the imaginary part of the result is then exactly the rounding
 * noise of a real quantity); 0 forces the generic path.
 * "lanczos_probe" (0 off, 1 on, 2 auto): after the k lowest pairs converged, lock them and probe
 * the complement with a fresh random direction so that degenerate copies are found. */
int psum_set_option(psum_handle_t h, const char* key, int64_t value);

/* info[0]=n_qubits [1]=merged terms [2]=distinct x-masks (G_x) [3]=passes [4]=groups read from
 * global memory ("remote") [5]=is_hermitian [6]=is_diagonal [7]=local qubits [8]=sum of patterns
 * [9]=component terms [10]=tile_bits [11]=distinct non-zero global x-parts [12]=flat groups
 * [13]=group records (x-groups after class splitting) [14]=group records served by a fast variant
 * [15]=partner load sets (bundles + singles; <= group records) [16]=quadratic groups (separable
 * table evaluation) */
int psum_info(psum_handle_t h, int64_t* info, int n_info);
/* Free-bit set (bit mask over local qubits) and group count of pass p of global part g. */
int psum_pass_info(psum_handle_t h, int g, int p, uint64_t* free_mask, int64_t* n_groups,
                   int64_t* n_remote_groups, int64_t* n_patterns);
/* Plan introspection, host only (no device needed): copies the device-facing arrays of pass p of
 * global part g exactly as the kernels read them, so the planner can be verified without a GPU
 * (tests/plan_interpreter.py).  counts: [0]=group records (48 B each) [1]=patterns [2]=component
 * terms [3]=chunks (32 B each) [4]=tile bits L [5]=non-free local bits [6]=has imaginary patterns
 * [7]=highest local qubit the pass combines [8]=bundle records (16 B each) [9]=register bits RB
 * [10]=quadratic groups (entries of quad_idx) [11]=1 if the pass runs as one big CTA per SM.
 * Any buffer may be NULL (call once for the counts); pat_tptr holds patterns+1 entries; chunk
 * records are 64 B.  Record layouts: qiskit_aqua_b200/csrc/psum_plan.h. */
int psum_plan_export(psum_handle_t h, int g, int p, int64_t counts[12], void* groups,
                     uint16_t* pat_zt, uint16_t* pat_inl, uint32_t* pat_tptr, uint64_t* term_zb,
                     double* term_val, void* chunks, uint8_t fbit[16], uint8_t nbit[64], void* bundles,
                     uint16_t* quad_idx);

/* Lower bounds of one H.psi on this rank computed from the plan (host only): out[0] = DRAM bytes
 * (store != 0: y is written), out[1] = shared-memory wavefronts, out[2] = group records, out[3] =
 * partner load sets.  bench.py divides them by the measured HBM bandwidth and by the shared-memory
 * pipe rate (1 wavefront per clock per SM) to get the roofline floors. */
int psum_plan_floors(psum_handle_t h, int store, double out[4]);

/* ---- H.psi and <psi|H|psi> ------------------------------------------------------------ */
/* y = H psi ; if expect_out != NULL also expect_out = {Re,Im} of sum_i conj(psi_i) y_i (host
 * pointer, forces a stream sync).  Replaces `mat_op._matrix.dot(psi)` (+ np.vdot) at
 * legacy/weighted_pauli_operator.py:621-622 and MatrixOp.eval primitive_ops/matrix_op.py:207. */
int psum_apply(psum_handle_t h, const void* psi_dev, void* y_dev, double* expect_out,
               psum_stream_t stream);
/* <psi|H|psi> without storing y.  Replaces evaluate_with_statevector
 * (legacy/weighted_pauli_operator.py:605-623), OperatorStateFn.eval
 * (state_fns/operator_state_fn.py:179-208) and _eval_aux_operators
 * (algorithms/eigen_solvers/numpy_eigen_solver.py:208-226). */
int psum_expect(psum_handle_t h, const void* psi_dev, double out[2], psum_stream_t stream);
/* The same for a state in HOST memory (pinned for full speed): psi_dev is a caller-provided device
 * buffer of 2^n amplitudes that receives the copy.  The H2D transfer is cut into chunks and the
 * passes that only combine amplitudes of one chunk run while later chunks are still in flight --
 * the end-to-end form of evaluate_with_statevector(numpy state). */
int psum_expect_host(psum_handle_t h, const void* psi_host, void* psi_dev, double out[2],
                     psum_stream_t stream);
/* <phi|H|psi> for a different bra (ListOp / transition elements). */
int psum_bra_apply(psum_handle_t h, const void* phi_dev, const void* psi_dev, double out[2],
                   psum_stream_t stream);
/* The same for `batch` state vectors stored back to back (VQE max_evals_grouped batching,
 * vqe.py:518, matrix_op.py:199-201).  out = 2*batch doubles on the host. */
int psum_expect_batch(psum_handle_t h, int batch, const void* psi_dev, double* out,
                      psum_stream_t stream);

/* Dense 2^n x 2^n matrix of the sum, row-major complex128 on the device (n <= 14): what the reference
 * obtains from to_matrix() (list_ops/list_op.py:306-329, primitive_ops/pauli_op.py:151-153) and feeds
 * to np.linalg.eig when k >= 2^n - 1 (numpy_eigen_solver.py:171-173).  Not on the hot path. */
int psum_to_dense(psum_handle_t h, void* mat_dev, psum_stream_t stream);

/* Inverse transform of psum_to_dense ("next" row 4 of SURVEY 8f): Pauli weights of a dense matrix,
 * replacing the 4^n sparse products tr(P M)/2^n of op_converter.to_weighted_pauli_operator
 * (legacy/op_converter.py:35-100) by one Walsh-Hadamard transform per x-mask on the device.
 * w_dev[x * 2^n + z] = weight(x, z) * (-i)^{popc(x & z)}; n_qubits <= 12. */
int psum_pauli_decompose(const void* mat_dev, int n_qubits, void* w_dev, psum_stream_t stream);
/* <psi|H_j|psi> of n_ops operators against one device state (aux operators
 * numpy_eigen_solver.py:208-226, qEOM commutators q_equation_of_motion.py:400-416): all passes are
 * queued back to back, one host synchronisation for the n_ops results (out = 2 * n_ops doubles). */
int psum_expect_multi(const psum_handle_t* hs, int n_ops, const void* psi_dev, double* out,
                      psum_stream_t stream);

/* ---- diagonal (all-Z) operators: numpy_eigen_solver.py:163-169 ------------------------ */
/* d_dev[i] = Re sum_k c_k (-1)^popc(i & z_k) as float64[N] on the device. */
int psum_diag(psum_handle_t h, double* d_dev, psum_stream_t stream);
/* k smallest diagonal entries, ascending, ties broken by lower index (np.sort / np.argsort). */
int psum_diag_kmin(psum_handle_t h, int k, double* vals, uint64_t* idx, psum_stream_t stream);

/* ---- eigensolver: replaces scipy eigs(H_csr, k, which='SR') numpy_eigen_solver.py:175-176 */
/* Device-resident Lanczos (full re-orthogonalisation while the basis fits in `mem_budget_bytes`
 * of device memory, otherwise 3-vector recurrence + a second pass for the vectors).
 * evals: k doubles (ascending).  evecs_dev: k*N complex128 on the device or NULL.
 * resid_out: k residual norms |H v - lambda v| or NULL.  Returns PSUM_ERR_NOT_HERMITIAN for
 * complex label coefficients. */
int psum_lanczos(psum_handle_t h, int k, double tol, int max_iter, uint64_t seed,
                 int64_t mem_budget_bytes, double* evals, void* evecs_dev, double* resid_out,
                 int* iters_out, psum_stream_t stream);

/* The same entry point for NON-Hermitian sums (complex label coefficients), which the reference's
 * eigs(which='SR') -- ARPACK's Arnoldi -- accepts at numpy_eigen_solver.py:175-176: Krylov-Schur
 * Arnoldi on the device matvec (sharded like psum_lanczos).  evals_re_im: 2k doubles (re, im
 * interleaved), ascending by (real, imaginary) part like the reference's eigval.argsort() (:177-180).
 * evecs_dev: k*N_loc complex128 unit-norm right eigenvectors or NULL; resid_out: k norms
 * |H v - lambda v| or NULL.  Works for Hermitian sums too (then it equals psum_lanczos). */
int psum_arnoldi(psum_handle_t h, int k, double tol, int max_iter, uint64_t seed,
                 int64_t mem_budget_bytes, double* evals_re_im, void* evecs_dev, double* resid_out,
                 int* iters_out, psum_stream_t stream);

/* ---- time evolution on the matvec: out = exp(-i t H) psi ("next" row of SURVEY 8f) ----------
 * Replaces the exact branch of MatrixOperator.evolve (legacy/matrix_operator.py:354-355) and
 * MatrixEvolution (evolutions/matrix_evolution.py:32-63) with Krylov (Lanczos) exponentiation:
 * krylov_dim basis vectors (0 = auto), adaptive sub-steps with local error <= tol.  out_dev may
 * equal psi_dev.  Hermitian sums only. */
int psum_evolve(psum_handle_t h, const void* psi_dev, void* out_dev, double t, double tol,
                int krylov_dim, int* steps_out, int* matvecs_out, psum_stream_t stream);

/* ---- ansatz -> psi on the device ("next" row 1 of SURVEY 8f; single GPU) -------------------
 * The step right before the hot path in VQE: the reference obtains psi from a simulator backend
 * (converters/circuit_sampler.py:330-333) and copies 2^n amplitudes per evaluation.  These three
 * calls let a host-side circuit description prepare psi in place on the device instead. */
int psum_state_basis(void* psi_dev, int n_qubits, uint64_t index, psum_stream_t stream);
/* u_re_im: row-major 2x2 complex matrix {u00, u01, u10, u11} as 8 doubles, applied to qubit q. */
int psum_gate_1q(void* psi_dev, int n_qubits, int q, const double* u_re_im, psum_stream_t stream);
/* kind 0 = CX(control, target), 1 = CZ(control, target). */
int psum_gate_2q(void* psi_dev, int n_qubits, int kind, int control, int target, psum_stream_t stream);

/* A block of gates whose qubits fit one shared-memory tile (n_free <= 12 qubits, ascending; 0, 1, 2
 * first when n_qubits >= 3): ONE sweep of the state for the whole block instead of one per gate.
 * gates_host: n_gates records of 80 bytes {int32 kind (0 = 2x2 unitary on a, 1 = CX control b target
 * a, 2 = CZ), int32 a, int32 b, int32 pad, double u[8]} with a, b positions inside free_qubits. */
int psum_gate_block(void* psi_dev, int n_qubits, const int32_t* free_qubits, int n_free,
                    const void* gates_host, int n_gates, psum_stream_t stream);

/* ---- vector helpers used by the host-side callers (device pointers, complex128) -------- */
/* Counter-based synthetic state: psi[i] = gauss(seed, first_index + i); its CPU twin is
 * oracle/psum_oracle.c:psum_oracle_state.  Not normalised. */
int psum_fill_state(void* psi_dev, uint64_t n_amps, uint64_t first_index, uint64_t seed,
                    psum_stream_t stream);
int psum_vec_dot(const void* a_dev, const void* b_dev, uint64_t n_amps, double out[2],
                 psum_stream_t stream);          /* sum conj(a) b */
int psum_vec_scale(void* a_dev, uint64_t n_amps, double s_re, double s_im, psum_stream_t stream);
int psum_vec_axpy(void* y_dev, const void* x_dev, uint64_t n_amps, double a_re, double a_im,
                  psum_stream_t stream);         /* y += a x */

/* ---- sharding over the top log2(world) qubits (one process per GPU) -------------------- */
/* 128-byte NCCL unique id (rank 0 creates it; the host broadcasts it, e.g. torch.distributed). */
int psum_nccl_unique_id(void* out128);
/* Re-plans the handle for this rank's shard.  unique_id128 == NULL -> no communicator is made
 * (plan inspection / CPU-side tests / caller-driven exchange through psum_apply_part). */
int psum_shard_init(psum_handle_t h, int rank, int world, const void* unique_id128);
/* Contribution of the terms whose global x-part is g (0 <= g < world) to this rank's y, reading
 * partner amplitudes from src_dev (= the shard of rank r^g; g == 0 -> the local shard).
 * accumulate != 0 adds into y_dev.  expect_out (host, 2 doubles, may be NULL), bra_dev = this
 * rank's own shard: with y_dev it receives sum conj(bra) y of y AFTER this call (so the call for
 * the last part yields the total); with y_dev == NULL it receives this part's contribution. */
int psum_apply_part(psum_handle_t h, int g, const void* src_dev, const void* bra_dev, void* y_dev,
                    int accumulate, double* expect_out, psum_stream_t stream);
/* Component terms this rank applies for global part g (rank sign folded in): for CPU-side
 * checks of the sharding algebra.  Arrays of capacity cap; returns the count in *n_out. */
int psum_shard_terms(psum_handle_t h, int g, int64_t cap, uint64_t* x_local, uint64_t* z_local,
                     double* coeff_re_im, int64_t* n_out);
/* Sharded H.psi with the NCCL amplitude exchange inside (needs psum_shard_init with an id):
 * recv_dev is a caller-provided exchange buffer of recv_amps >= N_loc amplitudes (may be NULL
 * when no term has a global x-part); it is used as a ring of floor(recv_amps / N_loc) <= 8
 * shard buffers, so that many exchanges are in flight ahead of the part being computed.  y_dev may be NULL (expectation only).  expect_out (may be NULL) receives
 * the all-reduced <psi|H|psi>. */
int psum_apply_sharded(psum_handle_t h, const void* psi_dev, void* y_dev, void* recv_dev,
                       uint64_t recv_amps, double* expect_out, psum_stream_t stream);
/* Sharded <psi|H|psi> for a shard that lives in HOST memory (the multi-GPU form of
 * psum_expect_host): psi_dev receives the copy; the chunk-local passes of the local part follow the
 * H2D copy chunk by chunk and the NCCL exchanges start when the last chunk has landed. */
int psum_expect_sharded_host(psum_handle_t h, const void* psi_host, void* psi_dev, void* recv_dev,
                             uint64_t recv_amps, double out[2], psum_stream_t stream);
/* Instrumentation of the last psum_apply_sharded / psum_expect_sharded_host call made with option
 * "profile_exchange" = 1: out[0] = exchanges, out[1] = bytes sent (= received) per exchange, out[2] =
 * sum of the exchange intervals on the communication stream in ms (CUDA events), out[3] = span from
 * the first post to the last arrival in ms.  Option "skip_exchange" = 1 issues the same launches
 * without the NCCL calls (compute-only timing; results are meaningless). */
int psum_exchange_stats(psum_handle_t h, double out[4]);
/* Host-side eigen-solver test hook (no device): which = 0 dense Jacobi (A = n*n), 1 tridiagonal
 * QL (A = d[n], e[n]).  w[n] ascending, V n*n eigenvectors in columns (may be NULL). */
int psum_test_eigh(int which, int n, const double* A, double* w, double* V);
/* Host-side test hooks of the Arnoldi path (no device): complex Schur form A = Z T Z^H of a dense n x n
 * matrix (row-major, re/im interleaved) with the diagonal ordered by (real, imaginary) part; and the
 * Krylov-Schur driver of psum_arnoldi run on a dense matrix with host vectors (m = basis size, 0 auto). */
int psum_test_schur(int n, const double* A_re_im, double* T_re_im, double* Z_re_im);
int psum_test_arnoldi(int n, const double* A_re_im, int k, int m, double tol, int max_iter, uint64_t seed,
                      double* evals_re_im, double* evecs_re_im, double* resid_out, int* iters_out);
int psum_allreduce_sum(psum_handle_t h, double* host_vals, int n); /* small host scalars */

#ifdef __cplusplus
}
#endif
#endif /* PSUM_H */

// psum_eig.h -- tiny host-side eigen-solvers for the projected (Lanczos / Arnoldi) matrices.
// The projected problem is at most ~128 x 128 (thick-restart basis) or a tridiagonal of a few
// hundred rows (3-vector recurrence); the 2^n-dimensional work stays on the device.
#pragma once
#include <math.h>

#include <algorithm>
#include <complex>
#include <numeric>
#include <vector>

namespace psum {

// Cyclic Jacobi for a dense real symmetric matrix A (n x n, row-major, destroyed).
// On return w = eigenvalues ascending, V (n x n row-major) has the eigenvectors in its columns.
inline void jacobi_eigh(int n, std::vector<double>& A, std::vector<double>& w, std::vector<double>& V) {
  V.assign((size_t)n * n, 0.0);
  for (int i = 0; i < n; ++i) V[(size_t)i * n + i] = 1.0;
  auto a = [&](int i, int j) -> double& { return A[(size_t)i * n + j]; };
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0.0, diag = 0.0;
    for (int p = 0; p < n; ++p) {
      diag += a(p, p) * a(p, p);
      for (int q = p + 1; q < n; ++q) off += a(p, q) * a(p, q);
    }
    if (off == 0.0 || sqrt(off) <= 1e-17 * sqrt(diag + off)) break;
    for (int p = 0; p < n - 1; ++p)
      for (int q = p + 1; q < n; ++q) {
        const double apq = a(p, q);
        if (apq == 0.0) continue;
        const double app = a(p, p), aqq = a(q, q);
        if (fabs(apq) < 1e-300) { a(p, q) = a(q, p) = 0.0; continue; }
        const double theta = (aqq - app) / (2.0 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        a(p, p) = app - t * apq;
        a(q, q) = aqq + t * apq;
        a(p, q) = a(q, p) = 0.0;
        for (int k = 0; k < n; ++k) {
          if (k != p && k != q) {
            const double akp = a(k, p), akq = a(k, q);
            a(k, p) = a(p, k) = c * akp - s * akq;
            a(k, q) = a(q, k) = s * akp + c * akq;
          }
          const double vkp = V[(size_t)k * n + p], vkq = V[(size_t)k * n + q];
          V[(size_t)k * n + p] = c * vkp - s * vkq;
          V[(size_t)k * n + q] = s * vkp + c * vkq;
        }
      }
  }
  std::vector<int> order(n);
  std::iota(order.begin(), order.end(), 0);
  std::sort(order.begin(), order.end(), [&](int i, int j) { return a(i, i) < a(j, j); });
  w.resize(n);
  std::vector<double> Vs((size_t)n * n);
  for (int c = 0; c < n; ++c) {
    w[c] = a(order[c], order[c]);
    for (int r = 0; r < n; ++r) Vs[(size_t)r * n + c] = V[(size_t)r * n + order[c]];
  }
  V.swap(Vs);
}

// Implicit-shift QL for a symmetric tridiagonal (d = diagonal [n], e = sub-diagonal [n], e[n-1]
// unused).  z (n x n row-major, identity on entry) receives the eigenvectors in its columns; pass
// an empty vector to skip them.  Eigenvalues are returned in d (unsorted).  Returns false if an
// eigenvalue needed more than 200 iterations.
inline bool tridiag_ql(int n, std::vector<double>& d, std::vector<double>& e, std::vector<double>& z) {
  const bool want_z = !z.empty();
  for (int l = 0; l < n; ++l) {
    int iter = 0, m;
    do {
      for (m = l; m < n - 1; ++m) {
        const double dd = fabs(d[m]) + fabs(d[m + 1]);
        if (fabs(e[m]) <= 2.3e-16 * dd) break;
      }
      if (m != l) {
        if (++iter > 200) return false;
        double g = (d[l + 1] - d[l]) / (2.0 * e[l]);
        double r = hypot(g, 1.0);
        g = d[m] - d[l] + e[l] / (g + (g >= 0 ? fabs(r) : -fabs(r)));
        double s = 1.0, c = 1.0, p = 0.0;
        int i;
        for (i = m - 1; i >= l; --i) {
          double f = s * e[i], b = c * e[i];
          r = hypot(f, g);
          e[i + 1] = r;
          if (r == 0.0) {
            d[i + 1] -= p;
            e[m] = 0.0;
            break;
          }
          s = f / r;
          c = g / r;
          g = d[i + 1] - p;
          r = (d[i] - g) * s + 2.0 * c * b;
          p = s * r;
          d[i + 1] = g + p;
          g = c * r - b;
          if (want_z)
            for (int k = 0; k < n; ++k) {
              f = z[(size_t)k * n + i + 1];
              z[(size_t)k * n + i + 1] = s * z[(size_t)k * n + i] + c * f;
              z[(size_t)k * n + i] = c * z[(size_t)k * n + i] - s * f;
            }
        }
        if (r == 0.0 && i >= l) continue;
        d[l] -= p;
        e[l] = g;
        e[m] = 0.0;
      }
    } while (m != l);
  }
  return true;
}

// ---------------------------------------------------------------------------------------------
// Complex Schur form of a small general matrix (the Rayleigh quotient of the Arnoldi basis).
//   A (n x n, row-major) is overwritten by the upper-triangular T, Z (n x n) receives the unitary
//   factor:  A = Z T Z^H.  Householder reduction to Hessenberg form, then the explicitly shifted
//   complex QR iteration with Givens rotations (Wilkinson shift, exceptional shifts every 10
//   iterations).  Returns false if an eigenvalue needed more than 60 iterations.
// ---------------------------------------------------------------------------------------------
typedef std::complex<double> cplx;

// rotation [c s; -conj(s) c] (c real) that maps (a, b) to (r, 0)
inline void givens_c(cplx a, cplx b, double& c, cplx& s) {
  const double na = std::abs(a), nb = std::abs(b);
  if (nb == 0.0) { c = 1.0; s = 0.0; return; }
  if (na == 0.0) { c = 0.0; s = std::conj(b) / nb; return; }
  const double nrm = hypot(na, nb);
  c = na / nrm;
  s = (a / na) * std::conj(b) / nrm;
}
// rows p, q of M (columns j0 .. n-1)  <-  G * rows
inline void rot_rows(int n, std::vector<cplx>& M, int p, int q, double c, cplx s, int j0) {
  for (int j = j0; j < n; ++j) {
    const cplx x = M[(size_t)p * n + j], y = M[(size_t)q * n + j];
    M[(size_t)p * n + j] = c * x + s * y;
    M[(size_t)q * n + j] = c * y - std::conj(s) * x;
  }
}
// columns p, q of M (rows 0 .. i1-1)  <-  columns * G^H
inline void rot_cols(int n, std::vector<cplx>& M, int p, int q, double c, cplx s, int i1) {
  for (int i = 0; i < i1; ++i) {
    const cplx x = M[(size_t)i * n + p], y = M[(size_t)i * n + q];
    M[(size_t)i * n + p] = c * x + std::conj(s) * y;
    M[(size_t)i * n + q] = c * y - s * x;
  }
}

inline bool complex_schur(int n, std::vector<cplx>& A, std::vector<cplx>& Z) {
  auto a = [&](int i, int j) -> cplx& { return A[(size_t)i * n + j]; };
  Z.assign((size_t)n * n, cplx(0.0, 0.0));
  for (int i = 0; i < n; ++i) Z[(size_t)i * n + i] = 1.0;
  // Hessenberg reduction: reflectors H_k = I - 2 v v^H on rows/columns k+1 .. n-1
  std::vector<cplx> v(n);
  for (int k = 0; k + 2 < n; ++k) {
    double nrm2 = 0.0;
    for (int i = k + 1; i < n; ++i) nrm2 += std::norm(a(i, k));
    const double tail2 = nrm2 - std::norm(a(k + 1, k));
    if (tail2 <= 0.0) continue;
    const double nrm = sqrt(nrm2);
    const cplx x0 = a(k + 1, k);
    const cplx ph = std::abs(x0) > 0.0 ? x0 / std::abs(x0) : cplx(1.0, 0.0);
    for (int i = 0; i < n; ++i) v[i] = 0.0;
    v[k + 1] = x0 + ph * nrm;
    for (int i = k + 2; i < n; ++i) v[i] = a(i, k);
    double vn2 = 0.0;
    for (int i = k + 1; i < n; ++i) vn2 += std::norm(v[i]);
    if (vn2 == 0.0) continue;
    const double sc = 2.0 / vn2;
    for (int j = 0; j < n; ++j) {          // A <- H A
      cplx d = 0.0;
      for (int i = k + 1; i < n; ++i) d += std::conj(v[i]) * a(i, j);
      d *= sc;
      for (int i = k + 1; i < n; ++i) a(i, j) -= v[i] * d;
    }
    for (int i = 0; i < n; ++i) {          // A <- A H ,  Z <- Z H
      cplx d = 0.0, e = 0.0;
      for (int j = k + 1; j < n; ++j) {
        d += a(i, j) * v[j];
        e += Z[(size_t)i * n + j] * v[j];
      }
      d *= sc;
      e *= sc;
      for (int j = k + 1; j < n; ++j) {
        a(i, j) -= d * std::conj(v[j]);
        Z[(size_t)i * n + j] -= e * std::conj(v[j]);
      }
    }
    for (int i = k + 2; i < n; ++i) a(i, k) = 0.0;
  }
  // shifted QR on the active window [l, en]
  std::vector<double> cs(n);
  std::vector<cplx> sn(n);
  int en = n - 1, iter = 0;
  while (en >= 0) {
    int l = en;
    while (l > 0) {
      const double tst = std::abs(a(l - 1, l - 1)) + std::abs(a(l, l));
      if (std::abs(a(l, l - 1)) <= 2.3e-16 * (tst > 0.0 ? tst : 1.0)) break;
      --l;
    }
    if (l > 0) a(l, l - 1) = 0.0;
    if (l == en) {
      --en;
      iter = 0;
      continue;
    }
    if (++iter > 60) return false;
    cplx shift;
    if (iter % 10 == 0) {
      shift = a(en, en) + cplx(std::abs(a(en, en - 1)), std::abs(a(en - 1, en - 1)) * 0.5);   // exceptional
    } else {
      const cplx p = a(en - 1, en - 1), q = a(en - 1, en), r = a(en, en - 1), t = a(en, en);
      const cplx half = 0.5 * (p - t), disc = std::sqrt(half * half + q * r);
      const cplx e1 = t + half + disc, e2 = t + half - disc;
      shift = std::abs(e1 - t) < std::abs(e2 - t) ? e1 : e2;
    }
    for (int i = l; i <= en; ++i) a(i, i) -= shift;
    for (int k = l; k < en; ++k) {     // R = Q^H (W - shift): zero the sub-diagonal
      givens_c(a(k, k), a(k + 1, k), cs[k], sn[k]);
      rot_rows(n, A, k, k + 1, cs[k], sn[k], k);
      a(k + 1, k) = 0.0;
    }
    for (int k = l; k < en; ++k) {     // R Q (and the rows above the window), Z Q
      rot_cols(n, A, k, k + 1, cs[k], sn[k], std::min(k + 2, en + 1));
      rot_cols(n, Z, k, k + 1, cs[k], sn[k], n);
    }
    for (int i = l; i <= en; ++i) a(i, i) += shift;
  }
  for (int i = 1; i < n; ++i)
    for (int j = 0; j < i; ++j) a(i, j) = 0.0;
  return true;
}

// Reorder a complex Schur form so that the diagonal ascends by (real part, imaginary part) -- the
// "smallest real part first" order of eigs(which='SR').  Adjacent entries are swapped by a rotation.
inline void schur_sort_sr(int n, std::vector<cplx>& T, std::vector<cplx>& Z) {
  auto t = [&](int i, int j) -> cplx& { return T[(size_t)i * n + j]; };
  auto less = [](cplx x, cplx y) { return x.real() < y.real() || (x.real() == y.real() && x.imag() < y.imag()); };
  for (int pass = 0; pass < n; ++pass) {
    bool any = false;
    for (int i = 0; i + 1 < n - pass; ++i) {
      if (!less(t(i + 1, i + 1), t(i, i))) continue;
      const cplx t11 = t(i, i), t22 = t(i + 1, i + 1);
      double c;
      cplx s;
      givens_c(t(i, i + 1), t22 - t11, c, s);
      rot_rows(n, T, i, i + 1, c, s, i);
      rot_cols(n, T, i, i + 1, c, s, i + 2);
      rot_cols(n, Z, i, i + 1, c, s, n);
      t(i + 1, i) = 0.0;
      t(i, i) = t22;
      t(i + 1, i + 1) = t11;
      any = true;
    }
    if (!any) break;
  }
}

// Eigenvector of the upper-triangular T for its i-th diagonal entry (back substitution), unit norm.
inline void triangular_eigvec(int n, const std::vector<cplx>& T, int i, std::vector<cplx>& y) {
  y.assign(n, cplx(0.0, 0.0));
  y[i] = 1.0;
  double tn = 0.0;
  for (int r = 0; r <= i; ++r)
    for (int c = r; c <= i; ++c) tn = std::max(tn, std::abs(T[(size_t)r * n + c]));
  const double tiny = std::max(tn, 1e-300) * 2.3e-16;
  for (int r = i - 1; r >= 0; --r) {
    cplx acc = 0.0;
    for (int c = r + 1; c <= i; ++c) acc += T[(size_t)r * n + c] * y[c];
    cplx den = T[(size_t)r * n + r] - T[(size_t)i * n + i];
    if (std::abs(den) < tiny) den = tiny;
    y[r] = -acc / den;
  }
  double nn = 0.0;
  for (int r = 0; r <= i; ++r) nn += std::norm(y[r]);
  nn = sqrt(nn);
  for (int r = 0; r <= i; ++r) y[r] /= nn;
}

}  // namespace psum

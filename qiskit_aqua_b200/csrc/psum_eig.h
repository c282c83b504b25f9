// psum_eig.h -- tiny host-side symmetric eigen-solvers for the projected (Lanczos) matrices.
// The projected problem is at most ~128 x 128 (thick-restart basis) or a tridiagonal of a few
// hundred rows (3-vector recurrence); the 2^n-dimensional work stays on the device.
#pragma once
#include <math.h>

#include <algorithm>
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

}  // namespace psum

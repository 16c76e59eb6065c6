// Host stand-in for the PC the reference selects in code: PCSetType(pc, PCILU); PCFactorSetLevels(pc, 2) on the finite-
// difference matrix P (elliptic.C:183-184) - and for PETSc's default PC of a SeqAIJ matrix, ILU(0), which the Stokes inner
// solves get on MatVVPC unless -vel_pc_type says otherwise.  The PC is PETSc's own and OUT OF SCOPE of the B200 path (BASELINE
// north_star: "stay PETSc's own ... timed separately"); this file exists so the command-line drivers and the solver-level parity
// tests can run the reference's DEFAULT solver configuration end to end.  It is the textbook level-of-fill ILU(k): natural
// ordering, lev(i,j) = min over k of lev(i,k) + lev(k,j) + 1 kept while <= k, IKJ numeric factorisation without pivoting or
// shifts (PETSc's defaults), unit-lower / upper triangular solves.  Plain C++ on host CSR arrays; nothing here touches the GPU.
#include <algorithm>
#include <climits>
#include <cmath>
#include <string>
#include <vector>

#include "../../include/spectral_b200.h"

namespace sb200 {
void set_last_error(const std::string& msg);
}

struct sb200_host_ilu {
  int n = 0, levels = 0;
  std::vector<int> rowptr, colidx, diag;  // factor pattern (L strictly below, U from the diagonal), columns increasing
  std::vector<double> val;                // L (unit diagonal not stored) and U
  std::vector<int> a_rowptr, a_colidx;    // pattern of A (to scatter new values on refactor)
  std::vector<double> work;
};

namespace {

int fail(int code, const std::string& msg) {
  sb200::set_last_error(msg);
  return code;
}

void symbolic(sb200_host_ilu* f) {
  const int n = f->n, K = f->levels;
  std::vector<std::vector<int>> ucols(n), ulev(n);  // U parts (j > i) of the finished rows with their levels
  std::vector<int> lev(n, INT_MAX), cols;
  f->rowptr.assign(1, 0);
  f->colidx.clear();
  f->diag.assign(n, -1);
  for (int i = 0; i < n; i++) {
    cols.assign(f->a_colidx.begin() + f->a_rowptr[i], f->a_colidx.begin() + f->a_rowptr[i + 1]);
    if (!std::binary_search(cols.begin(), cols.end(), i)) cols.insert(std::lower_bound(cols.begin(), cols.end(), i), i);
    for (int c : cols) lev[c] = 0;
    for (size_t p = 0; p < cols.size() && cols[p] < i; p++) {  // pivots in increasing order; insertions land behind p
      const int k = cols[p], lik = lev[k];
      for (size_t q = 0; q < ucols[k].size(); q++) {
        const int j = ucols[k][q];
        const long long nl = (long long)lik + ulev[k][q] + 1;
        if (nl > K) continue;
        if (lev[j] == INT_MAX) cols.insert(std::lower_bound(cols.begin() + p + 1, cols.end(), j), j);
        lev[j] = std::min(lev[j], (int)nl);
      }
    }
    for (int c : cols) {
      if (c == i) f->diag[i] = (int)f->colidx.size();
      if (c > i) {
        ucols[i].push_back(c);
        ulev[i].push_back(lev[c]);
      }
      f->colidx.push_back(c);
      lev[c] = INT_MAX;
    }
    f->rowptr.push_back((int)f->colidx.size());
  }
}

int numeric(sb200_host_ilu* f, const double* avals) {
  const int n = f->n;
  std::vector<int> pos(n, -1);
  f->val.assign(f->colidx.size(), 0.0);
  for (int i = 0; i < n; i++) {
    const int r0 = f->rowptr[i], r1 = f->rowptr[i + 1];
    for (int p = r0; p < r1; p++) pos[f->colidx[p]] = p;
    for (int p = f->a_rowptr[i]; p < f->a_rowptr[i + 1]; p++) f->val[pos[f->a_colidx[p]]] = avals[p];
    for (int p = r0; p < f->diag[i]; p++) {
      const int k = f->colidx[p];
      const double l = f->val[p] / f->val[f->diag[k]];
      f->val[p] = l;
      for (int q = f->diag[k] + 1; q < f->rowptr[k + 1]; q++) {
        const int t = pos[f->colidx[q]];
        if (t >= 0) f->val[t] -= l * f->val[q];
      }
    }
    for (int p = r0; p < r1; p++) pos[f->colidx[p]] = -1;
    const double d = f->val[f->diag[i]];
    if (d == 0.0 || !std::isfinite(d)) return fail(SB200_ERR_USER, "ILU: zero pivot in row " + std::to_string(i));  // PETSc: "Zero pivot row"
  }
  return 0;
}

}  // namespace

extern "C" {

int sb200_host_ilu_create(int n, const int* rowptr, const int* colidx, const double* vals, int levels, sb200_host_ilu** out) {
  if (!rowptr || !colidx || !vals || !out) return fail(SB200_ERR_ARG, "null pointer");
  *out = nullptr;
  if (n < 1 || levels < 0) return fail(SB200_ERR_USER, "ILU: n >= 1 and levels >= 0 required");
  if (rowptr[0] != 0) return fail(SB200_ERR_USER, "ILU: rowptr[0] must be 0");
  for (int i = 0; i < n; i++) {
    if (rowptr[i + 1] < rowptr[i]) return fail(SB200_ERR_USER, "ILU: rowptr must not decrease");
    for (int p = rowptr[i]; p < rowptr[i + 1]; p++) {
      if (colidx[p] < 0 || colidx[p] >= n) return fail(SB200_ERR_USER, "ILU: column index out of range");
      if (p > rowptr[i] && colidx[p] <= colidx[p - 1]) return fail(SB200_ERR_USER, "ILU: columns must increase within a row");
    }
  }
  sb200_host_ilu* f = new sb200_host_ilu();
  f->n = n;
  f->levels = levels;
  f->a_rowptr.assign(rowptr, rowptr + n + 1);
  f->a_colidx.assign(colidx, colidx + rowptr[n]);
  f->work.resize(n);
  symbolic(f);
  const int rc = numeric(f, vals);
  if (rc) {
    delete f;
    return rc;
  }
  *out = f;
  return 0;
}

int sb200_host_ilu_refactor(sb200_host_ilu* f, const double* vals) {
  if (!f || !vals) return fail(SB200_ERR_ARG, "null pointer");
  return numeric(f, vals);
}

int sb200_host_ilu_solve(const sb200_host_ilu* f, const double* b, double* x) {
  if (!f || !b || !x) return fail(SB200_ERR_ARG, "null pointer");
  const int n = f->n;
  for (int i = 0; i < n; i++) {  // L y = b (unit diagonal)
    double s = b[i];
    for (int p = f->rowptr[i]; p < f->diag[i]; p++) s -= f->val[p] * x[f->colidx[p]];
    x[i] = s;
  }
  for (int i = n - 1; i >= 0; i--) {  // U x = y
    double s = x[i];
    for (int p = f->diag[i] + 1; p < f->rowptr[i + 1]; p++) s -= f->val[p] * x[f->colidx[p]];
    x[i] = s / f->val[f->diag[i]];
  }
  return 0;
}

int sb200_host_ilu_nnz(const sb200_host_ilu* f, long long* nnz) {
  if (!f || !nnz) return fail(SB200_ERR_ARG, "null pointer");
  *nnz = (long long)f->colidx.size();
  return 0;
}

int sb200_host_ilu_get(const sb200_host_ilu* f, int* rowptr, int* colidx, double* vals) {
  if (!f) return fail(SB200_ERR_ARG, "null pointer");
  if (rowptr) std::copy(f->rowptr.begin(), f->rowptr.end(), rowptr);
  if (colidx) std::copy(f->colidx.begin(), f->colidx.end(), colidx);
  if (vals) std::copy(f->val.begin(), f->val.end(), vals);
  return 0;
}

int sb200_host_ilu_destroy(sb200_host_ilu* f) {
  delete f;
  return 0;
}

}  // extern "C"

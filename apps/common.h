// Shared by the native executables apps/elliptic.cpp and apps/stokes.cpp: the slice of the PETSc options database they read,
// and the HOST stand-in for PETSc's PC on a finite-difference matrix (ILU(k) / Jacobi / none; the PC is PETSc's own and out of
// scope of the B200 path, see include/spectral_b200.h "host stand-in").
#pragma once
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../include/sb200_reference_api.h"
#include "../include/spectral_b200.h"

#define CHK(expr)                                                                          \
  do {                                                                                     \
    PetscErrorCode _e = (expr);                                                            \
    if (_e) {                                                                              \
      fprintf(stderr, "%s:%d error %d: %s\n", __FILE__, __LINE__, _e, sb200_last_error()); \
      return _e;                                                                           \
    }                                                                                      \
  } while (0)

namespace app {

struct Options {
  std::map<std::string, std::string> kv;
  std::map<std::string, bool> used;
  static bool number(const char* s) {
    char* end = nullptr;
    strtod(s, &end);
    return end != s && *end == 0;
  }
  int parse(int argc, char** argv) {
    for (int i = 1; i < argc;) {
      const char* a = argv[i];
      if (a[0] != '-' || number(a)) {
        fprintf(stderr, "error: expected an option name, got '%s'\n", a);
        return 83;
      }
      if (i + 1 < argc && (argv[i + 1][0] != '-' || number(argv[i + 1]))) {
        kv[a + 1] = argv[i + 1];
        i += 2;
      } else {
        kv[a + 1] = "";
        i += 1;
      }
    }
    return 0;
  }
  const std::string* get(const char* name) {
    used[name] = true;
    auto it = kv.find(name);
    return it == kv.end() ? nullptr : &it->second;
  }
  bool has(const char* name) { return get(name) != nullptr; }
  int integer(const char* name, int dflt) {
    const std::string* v = get(name);
    return v && !v->empty() ? atoi(v->c_str()) : dflt;
  }
  double real(const char* name, double dflt) {
    const std::string* v = get(name);
    return v && !v->empty() ? atof(v->c_str()) : dflt;
  }
  std::string str(const char* name, const char* dflt) {
    const std::string* v = get(name);
    return v && !v->empty() ? *v : std::string(dflt);
  }
  // PetscOptionsIntArray("-dim", ...) (elliptic.C:141, stokes.C:407): up to maxlen comma-separated extents; 0 = option absent
  int int_array(const char* name, int* out, int maxlen) {
    const std::string* v = get(name);
    if (!v) return 0;
    int n = 0;
    const char* p = v->c_str();
    while (*p) {
      char* end = nullptr;
      const long x = strtol(p, &end, 10);
      if (end == p) break;
      if (n == maxlen) return -1;
      out[n++] = (int)x;
      p = (*end == ',') ? end + 1 : end;
    }
    return n ? n : -1;
  }
  void warn_unused() const {  // PETSc's message at PetscFinalize
    for (auto& kvp : kv)
      if (!used.count(kvp.first)) printf("WARNING! There are options you set that were not used: -%s\n", kvp.first.c_str());
  }
};

struct HostPc {  // stand-in for PETSc's PC on P: ilu (levels), jacobi, none
  std::string type = "ilu";
  int levels = 0;
  sb200_host_ilu* ilu = nullptr;
  std::vector<int> rowptr, colidx;
  std::vector<double> vals, diag, hx, hy;
  double* d_diag = nullptr;  // jacobi: the diagonal of P on the device, so that applying it needs no host round trip
  int n = 0;
  ~HostPc() { release(); }
  void release() {  // call before main returns (device memory)
    if (ilu) sb200_host_ilu_destroy(ilu);
    if (d_diag) sb200_free(d_diag);
    ilu = nullptr;
    d_diag = nullptr;
  }
  static bool known(const std::string& t) { return t == "ilu" || t == "jacobi" || t == "none"; }
  // PCSetUp: bring the values of P down (pattern once) and (re)factor
  int setup(Mat P) {
    if (type == "none") return 0;
    PetscInt rows, nz;
    CHK(MatGetSize(P, &rows, PETSC_NULL));
    if (type == "jacobi") {  // the diagonal is taken on the device: the matrix never comes down
      n = rows;
      if (!d_diag) CHK(sb200_malloc((void**)&d_diag, (size_t)rows * sizeof(double)));
      CHK(sb200_csr_diagonal(rows, P->d_rowptr, P->d_colidx, P->d_vals, d_diag, nullptr));
      diag.resize(rows);  // host copy for apply_host (the -saddle_on_host cross-check)
      CHK(sb200_memcpy_d2h(diag.data(), d_diag, (size_t)rows * sizeof(double), nullptr));
      return sb200_stream_sync(nullptr);
    }
    CHK(MatSeqAIJGetCSRHost(P, &nz, PETSC_NULL, PETSC_NULL, PETSC_NULL));
    const bool first = rowptr.empty();
    if (first) {
      n = rows;
      rowptr.resize(rows + 1);
      colidx.resize(nz);
      vals.resize(nz);
      hx.resize(rows);
      hy.resize(rows);
      CHK(MatSeqAIJGetCSRHost(P, PETSC_NULL, rowptr.data(), colidx.data(), vals.data()));
    } else {
      CHK(MatSeqAIJGetCSRHost(P, PETSC_NULL, PETSC_NULL, PETSC_NULL, vals.data()));
    }
    if (type == "ilu") {
      if (first) CHK(sb200_host_ilu_create(rows, rowptr.data(), colidx.data(), vals.data(), levels, &ilu));
      else CHK(sb200_host_ilu_refactor(ilu, vals.data()));
    }
    return 0;
  }
  // z = M^-1 r on host arrays
  int apply_host(const double* r, double* z) const {
    if (type == "none") {
      std::memcpy(z, r, sizeof(double) * (size_t)n);
      return 0;
    }
    if (type == "jacobi") {
      for (int i = 0; i < n; i++) z[i] = r[i] / diag[i];  // the same division the device form performs
      return 0;
    }
    return sb200_host_ilu_solve(ilu, r, z);
  }
  // the same on device arrays: jacobi and none stay on the device (VecPointwiseDivide / copy); ilu goes down, applies, comes up
  int apply_device(const double* d_x, double* d_y, void* stream) {
    const size_t bytes = (size_t)n * sizeof(double);
    if (type == "jacobi") return sb200_vec_pointwise_divide(n, d_x, d_diag, d_y, stream);
    if (type == "none") return sb200_memcpy_d2d(d_y, d_x, bytes, stream);
    int rc = sb200_memcpy_d2h(hx.data(), d_x, bytes, stream);
    if (!rc) rc = sb200_stream_sync(stream);
    if (!rc) rc = apply_host(hx.data(), hy.data());
    if (!rc) rc = sb200_memcpy_h2d(d_y, hy.data(), bytes, stream);
    return rc ? rc : sb200_stream_sync(stream);  // hy is reused by the next application
  }
};

inline double norm2(const std::vector<double>& v) {
  double s = 0;
  for (double x : v) s += x * x;
  return sqrt(s);
}

inline double norm_inf(const std::vector<double>& v) {
  double s = 0;
  for (double x : v) s = fmax(s, fabs(x));
  return s;
}

}  // namespace app

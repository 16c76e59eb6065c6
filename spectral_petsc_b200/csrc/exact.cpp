// Manufactured solutions of the reference's drivers, evaluated on the host (set-up work, once per run):
//   CreateExactSolution        elliptic.C:594-677  (-exact 0 / 1 / 2, -cos_scale, -gamma, -exponent)
//   StokesCreateExactSolution  stokes.C:942-1003 with StokesExact0..3 (:1948-2034) and StokesDirichlet (:2039-2050)
// Outputs follow the reference's Vec layouts: interior values in walk order (elliptic: one value per interior node; Stokes:
// AoS [v_0..v_{d-1}, p]), boundary values in walk order (Stokes: d velocity components per boundary node).  Plain C++,
// no device needed; the results are uploaded with sb200_elliptic_set_dirichlet / _set_rhs, sb200_stokes_set_dirichlet / _set_force.
#include <cmath>
#include <string>

#include "../../include/spectral_b200.h"

namespace sb200 {
void set_last_error(const std::string& msg);
}

namespace {

const double kPi = 3.14159265358979323846;  // chebyshev.h:10 (PI) = PETSC_PI

// Walks the grid like BlockIt (util.C:8-40): last axis fastest; calls f(node coordinates, on boundary?).
template <class F>
void walk(int d, const int* dim, F f) {
  int ind[10] = {0};
  double x[10];
  long long m = 1;
  for (int j = 0; j < d; j++) m *= dim[j];
  for (long long node = 0; node < m; node++) {
    bool bdy = false;
    for (int j = 0; j < d; j++) {
      x[j] = cos(ind[j] * kPi / (dim[j] - 1));  // elliptic.C:279, stokes.C:297
      bdy = bdy || ind[j] == 0 || ind[j] == dim[j] - 1;
    }
    f(x, bdy);
    for (int j = d - 1; j >= 0; j--) {
      if (++ind[j] < dim[j]) break;
      ind[j] = 0;
    }
  }
}

int check_grid(int d, const int* dim, int maxd) {
  if (!dim) {
    sb200::set_last_error("null pointer");
    return SB200_ERR_ARG;
  }
  if (d < 1 || d > maxd) {
    sb200::set_last_error("dimension count out of range");
    return SB200_ERR_USER;
  }
  for (int j = 0; j < d; j++)
    if (dim[j] < 3) {
      sb200::set_last_error("each extent must be >= 3 (needs an interior node)");
      return SB200_ERR_USER;
    }
  return 0;
}

}  // namespace

extern "C" {

int sb200_elliptic_exact_solution(int d, const int* dim, int exact, double cos_scale, double gamma, double exponent, double* h_u,
                                  double* h_u2, double* h_dirichlet) {
  if (int rc = check_grid(d, dim, 10)) return rc;
  if (exact < 0 || exact > 2) {
    sb200::set_last_error("Choose an exact solution.");  // elliptic.C:657
    return SB200_ERR_USER;
  }
  double s = 0.5;
  if (exact == 0 || exact == 3) s *= cos_scale;  // elliptic.C:605-610
  long long gi = 0, di = 0;
  walk(d, dim, [&](const double* x, bool bdy) {
    double v = 1.0, w = 0.0;
    switch (exact) {
      case 0: {  // separable cosine, handles the nonlinearity (elliptic.C:620-632)
        for (int j = 0; j < d; j++) v *= cos(s * kPi * x[j]);
        const double eta = 1.0 + gamma * pow(v, exponent);
        const double deta = (fabs(exponent) < 1e-10) ? 0.0 : gamma * exponent * pow(v, exponent - 1.0);
        for (int j = 0; j < d; j++) {
          double dv = 1.0;
          for (int k = 0; k < d; k++) dv *= (k == j) ? -s * kPi * sin(s * kPi * x[k]) : cos(s * kPi * x[k]);
          const double d2v = -(s * kPi) * (s * kPi) * v;
          w += deta * dv * dv + eta * d2v;
        }
        w = -w;
      } break;
      case 1:  // separable quadratics, zero on the boundary (elliptic.C:633-643)
        for (int j = 0; j < d; j++) {
          v *= (1 - x[j]) * (1 + x[j]);
          double z = 1.0;
          for (int k = 0; k < d; k++)
            if (k != j) z *= 2.0 * (1 - x[k]) * (1 + x[k]);
          w += z;
        }
        break;
      case 2:  // separable polynomials, nonzero on the boundary (elliptic.C:644-655)
        for (int j = 0; j < d; j++) {
          v *= pow(x[j], 4 + j);
          double z = 1.0;
          for (int k = 0; k < d; k++) z *= (k == j) ? (4 + k) * (3 + k) * pow(x[k], 2 + k) : pow(x[k], 4 + k);
          w -= z;
        }
        break;
    }
    if (bdy) {
      if (h_dirichlet) h_dirichlet[di] = v;  // scatterLD (elliptic.C:672)
      di++;
    } else {
      if (h_u) h_u[gi] = v;    // scatterLG of w[0] (elliptic.C:668)
      if (h_u2) h_u2[gi] = w;  // scatterLG of w[1] (elliptic.C:670)
      gi++;
    }
  });
  return 0;
}

// One node of StokesExact0..3 (stokes.C:1948-2034): value = [u_0..u_{d-1}, p], rhs = the forcing, d + 1 numbers each (either
// may be NULL).  -exact 2 in three dimensions leaves value[3] unset in the reference (stokes.C:2004-2005); it is 0 here.
int sb200_stokes_exact_eval(int exact, int d, const double* c, double* value, double* rhs_out) {
  if (d < 2 || d > 3 || !c) {
    sb200::set_last_error("the Stokes problem needs 2 or 3 dimensions");
    return SB200_ERR_USER;
  }
  if (exact < 0 || exact > 3) {
    sb200::set_last_error("Exact solution not implemented");  // stokes.C:452
    return SB200_ERR_SUP;
  }
  if (exact == 3 && d != 2) {
    sb200::set_last_error("StokesExact3 only implemented for dimension 2");  // stokes.C:2022
    return SB200_ERR_USER;
  }
  double val[4] = {0, 0, 0, 0}, rhs[4] = {0, 0, 0, 0};
  if (exact == 1 || exact == 2) {  // StokesExact1 / 2 (stokes.C:1963-2012)
    const double eta = 1.0;
    const double u = sin(0.5 * kPi * c[0]) * cos(0.5 * kPi * c[1]);
    const double v = -cos(0.5 * kPi * c[0]) * sin(0.5 * kPi * c[1]);
    val[0] = u;
    val[1] = v;
    rhs[0] = (0.5 * kPi) * (0.5 * kPi) * eta * u;
    rhs[1] = (0.5 * kPi) * (0.5 * kPi) * eta * v;
    if (exact == 1) {
      val[d] = 0.25 * (cos(kPi * c[0]) + cos(kPi * c[1])) + 10 * (c[0] + c[1]);
      rhs[0] += -0.25 * kPi * sin(kPi * c[0]) + 10;
      rhs[1] += -0.25 * kPi * sin(kPi * c[1]) + 10;
    }
  } else if (exact == 3) {  // StokesExact3 (stokes.C:2016-2034): shear flow u = y + 1
    val[0] = c[1] + 1.0;
  }
  for (int k = 0; k <= d; k++) {
    if (value) value[k] = val[k];
    if (rhs_out) rhs_out[k] = rhs[k];
  }
  return 0;
}

int sb200_stokes_exact_solution(int d, const int* dim, int exact, double* h_u, double* h_u2, double* h_dirichlet) {
  if (int rc = check_grid(d, dim, 3)) return rc;
  if (d < 2) {
    sb200::set_last_error("the Stokes problem needs 2 or 3 dimensions");
    return SB200_ERR_USER;
  }
  if (exact < 0 || exact > 3) {
    sb200::set_last_error("Exact solution not implemented");  // stokes.C:452
    return SB200_ERR_SUP;
  }
  if (exact == 3 && d != 2) {
    sb200::set_last_error("StokesExact3 only implemented for dimension 2");  // stokes.C:2022
    return SB200_ERR_USER;
  }
  long long gi = 0, di = 0;
  walk(d, dim, [&](const double* c, bool bdy) {
    double val[4], rhs[4];
    sb200_stokes_exact_eval(exact, d, c, val, rhs);  // arguments validated above
    if (bdy) {  // StokesDirichlet evaluates the exact solution (stokes.C:2039-2050); d velocity dofs per boundary node (:796-801)
      if (h_dirichlet)
        for (int k = 0; k < d; k++) h_dirichlet[di * d + k] = val[k];
      di++;
    } else {
      if (h_u)
        for (int k = 0; k <= d; k++) h_u[gi * (d + 1) + k] = val[k];
      if (h_u2)
        for (int k = 0; k <= d; k++) h_u2[gi * (d + 1) + k] = rhs[k];
      gi++;
    }
  });
  return 0;
}

}  // extern "C"

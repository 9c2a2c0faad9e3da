// alpha.cpp -- host-side running coupling (O(nQ2) scalar work, kept on the host).
//
// Reference: f90src/qcd/alpha_qcd.f90:22-127 and constants.f90:16-39.  Two-loop beta
// function (norder = 1), 17 RK4 steps in ln Q^2 from the nearest flavour threshold,
// exact-equality shortcuts at MZ^2, mt^2, mb^2, mc^2.
#include <cmath>
#include "tkb_internal.h"

namespace tkb {
namespace {
const double kCF = 4.0 / 3.0, kCA = 3.0, kTF = 0.5, kPi = 3.141592653589793;
const double cMass2 = 1.29 * 1.29, bMass2 = 4.18 * 4.18, tMass2 = 172.44 * 172.44;
const double alphaMZ = 0.118, MZ2 = 91.1876 * 91.1876;

double beta_qcd(double alpha, int nfl)
{
    const double fn = (double)nfl;
    const double beta0 = (11.0 * kCA - 4.0 * fn * kTF) / (12.0 * kPi);
    const double beta1 = (17.0 * kCA * kCA - fn * kTF * (10.0 * kCA + 6.0 * kCF)) / (24.0 * (kPi * kPi));
    double beta = beta0 + alpha * beta1;
    return -beta * (alpha * alpha);
}

double evolve_fixed_nfl(double alpha0, double Q20, double Q2, int nfl)
{
    const int nsteps = 17;
    const double h = std::log(Q2 / Q20) / (double)nsteps;
    double alpha = alpha0;
    for (int i = 0; i < nsteps; i++) {
        const double k1 = beta_qcd(alpha, nfl);
        const double k2 = beta_qcd(alpha + h * k1 / 2.0, nfl);
        const double k3 = beta_qcd(alpha + h * k2 / 2.0, nfl);
        const double k4 = beta_qcd(alpha + h * k3, nfl);
        alpha = alpha + h / 6.0 * (k1 + 2.0 * k2 + 2.0 * k3 + k4);
    }
    return alpha;
}
double alpha_mt2() { return evolve_fixed_nfl(alphaMZ, MZ2, tMass2, 5); }
double alpha_mb2() { return evolve_fixed_nfl(alphaMZ, MZ2, bMass2, 5); }
double alpha_mc2() { return evolve_fixed_nfl(alpha_mb2(), bMass2, cMass2, 4); }
}  // namespace

double alpha_qcd(double Q2)
{
    if (Q2 == MZ2) return alphaMZ;
    if (Q2 == tMass2) return alpha_mt2();
    if (Q2 == bMass2) return alpha_mb2();
    if (Q2 == cMass2) return alpha_mc2();
    if (Q2 > tMass2) return evolve_fixed_nfl(alpha_mt2(), tMass2, Q2, 6);
    if (Q2 > bMass2) return evolve_fixed_nfl(alpha_mb2(), bMass2, Q2, 5);
    if (Q2 > cMass2) return evolve_fixed_nfl(alpha_mc2(), cMass2, Q2, 4);
    return evolve_fixed_nfl(alpha_mc2(), cMass2, Q2, 3);
}

int neff(double Q2)
{
    if (Q2 >= tMass2) return 6;
    if (Q2 >= bMass2) return 5;
    if (Q2 >= cMass2) return 4;
    return 3;
}
}  // namespace tkb

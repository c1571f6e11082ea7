// Piecewise-polynomial coefficient interpolation on the host.
//
// Restates what QuTiP 5 does with an array coefficient handed to QobjEvo
// together with `tlist` (reference call site:
// pulser-simulation/pulser_simulation/hamiltonian.py:436): order 3 = cubic
// interpolating spline with not-a-knot end conditions
// (scipy.interpolate.make_interp_spline(k=3) default), order 1 = linear,
// order 0 = previous value.  The interpolant is stored per sampling interval
// as a cubic in tau = t - t_i so that exact integrals (Magnus moments) over
// arbitrary [a, b] are cheap.
#pragma once
#include <algorithm>
#include <cmath>
#include <complex>
#include <vector>

namespace pb200 {

using cplx = std::complex<double>;

template <typename T>
struct PiecewiseCubic {
    // piece i on [x[i], x[i+1]]: c0 + c1*tau + c2*tau^2 + c3*tau^3
    std::vector<T> c0, c1, c2, c3;
    T y_last = T(0);  // last sample (needed by the order-0 rule at the end point)

    int pieces() const { return (int)c0.size(); }

    T eval_piece(int i, double tau) const {
        return c0[i] + tau * (c1[i] + tau * (c2[i] + tau * c3[i]));
    }
    // integral over [t0, t1] (local tau) of S and of (tau - m) * S
    void moments_piece(int i, double t0, double t1, double m, T& i0, T& i1) const {
        // antiderivatives in powers of tau
        auto P = [&](double t) {
            return t * (c0[i] + t * (c1[i] / 2.0 + t * (c2[i] / 3.0 + t * c3[i] / 4.0)));
        };
        auto Q = [&](double t) {  // integral of tau * S
            return t * t * (c0[i] / 2.0 + t * (c1[i] / 3.0 + t * (c2[i] / 4.0 + t * c3[i] / 5.0)));
        };
        T p = P(t1) - P(t0);
        T q = Q(t1) - Q(t0);
        i0 = p;
        i1 = q - m * p;
    }
};

// Not-a-knot cubic spline through (x_i, y_i), n >= 4 (n == 3: parabola,
// n == 2: line -- scipy lowers the degree the same way only when asked; the
// reference guarantees >= 4 points, simulation.py:183-186).
template <typename T>
PiecewiseCubic<T> make_interpolant(const double* x, const T* y, int n, int order) {
    PiecewiseCubic<T> pc;
    int np = n - 1;
    pc.c0.resize(np); pc.c1.assign(np, T(0)); pc.c2.assign(np, T(0)); pc.c3.assign(np, T(0));
    for (int i = 0; i < np; ++i) pc.c0[i] = y[i];
    pc.y_last = y[n - 1];
    if (order == 0 || n < 2) return pc;
    if (order == 1 || n < 4) {
        for (int i = 0; i < np; ++i) pc.c1[i] = (y[i + 1] - y[i]) / (x[i + 1] - x[i]);
        return pc;
    }
    // second derivatives M_0..M_{n-1}
    std::vector<double> h(np);
    for (int i = 0; i < np; ++i) h[i] = x[i + 1] - x[i];
    std::vector<T> M(n);
    // unknowns M_1..M_{n-2}; not-a-knot: M_0 = ((h0+h1) M_1 - h0 M_2)/h1,
    // M_{n-1} = ((h_{n-2}+h_{n-3}) M_{n-2} - h_{n-2} M_{n-3}) / h_{n-3}
    int m = n - 2;
    std::vector<double> lo(m, 0.0), di(m, 0.0), up(m, 0.0);
    std::vector<T> rhs(m);
    for (int k = 0; k < m; ++k) {
        int i = k + 1;
        lo[k] = h[i - 1];
        di[k] = 2.0 * (h[i - 1] + h[i]);
        up[k] = h[i];
        rhs[k] = 6.0 * ((y[i + 1] - y[i]) / h[i] - (y[i] - y[i - 1]) / h[i - 1]);
    }
    {   // fold M_0 into row 0
        double a = h[0];  // coefficient of M_0 in row i=1
        di[0] += a * (h[0] + h[1]) / h[1];
        up[0] -= a * h[0] / h[1];
        lo[0] = 0.0;
        // fold M_{n-1} into the last row
        double b = h[n - 2];
        di[m - 1] += b * (h[n - 2] + h[n - 3]) / h[n - 3];
        lo[m - 1] -= b * h[n - 2] / h[n - 3];
        up[m - 1] = 0.0;
    }
    // Thomas algorithm
    std::vector<double> cp(m);
    std::vector<T> dp(m);
    cp[0] = up[0] / di[0];
    dp[0] = rhs[0] / di[0];
    for (int k = 1; k < m; ++k) {
        double den = di[k] - lo[k] * cp[k - 1];
        cp[k] = up[k] / den;
        dp[k] = (rhs[k] - lo[k] * dp[k - 1]) / den;
    }
    M[m] = dp[m - 1];  // M index = k+1
    for (int k = m - 2; k >= 0; --k) M[k + 1] = dp[k] - cp[k] * M[k + 2];
    M[0] = ((h[0] + h[1]) * M[1] - h[0] * M[2]) / h[1];
    M[n - 1] = ((h[n - 2] + h[n - 3]) * M[n - 2] - h[n - 2] * M[n - 3]) / h[n - 3];
    for (int i = 0; i < np; ++i) {
        pc.c1[i] = (y[i + 1] - y[i]) / h[i] - h[i] * (2.0 * M[i] + M[i + 1]) / 6.0;
        pc.c2[i] = M[i] / 2.0;
        pc.c3[i] = (M[i + 1] - M[i]) / (6.0 * h[i]);
    }
    return pc;
}

// locate the piece containing t (clamped)
inline int find_piece(const std::vector<double>& x, double t) {
    int n = (int)x.size();
    if (t <= x[0]) return 0;
    if (t >= x[n - 1]) return n - 2;
    int i = (int)(std::upper_bound(x.begin(), x.end(), t) - x.begin()) - 1;
    return std::min(std::max(i, 0), n - 2);
}

template <typename T>
T eval_at(const PiecewiseCubic<T>& pc, const std::vector<double>& x, double t, int order) {
    t = std::min(std::max(t, x.front()), x.back());
    int i = find_piece(x, t);
    if (order == 0 && t >= x.back()) return pc.y_last;  // the last sample holds only at the end point
    return pc.eval_piece(i, t - x[i]);
}

// B0 = int_a^b S dt,  B1 = (1/(b-a)) int_a^b (t - (a+b)/2) S dt
template <typename T>
void magnus_moments(const PiecewiseCubic<T>& pc, const std::vector<double>& x,
                    double a, double b, T& B0, T& B1) {
    B0 = T(0); B1 = T(0);
    if (!(b > a)) return;
    double tm = 0.5 * (a + b);
    int ia = find_piece(x, a), ib = find_piece(x, b);
    if (b <= x[ib] && ib > ia) ib -= 1;  // b exactly on a knot: stop in the piece before
    for (int i = ia; i <= ib; ++i) {
        double t0 = std::max(a, x[i]) - x[i];
        double t1 = std::min(b, x[i + 1]) - x[i];
        if (t1 <= t0) continue;
        T i0, i1;
        pc.moments_piece(i, t0, t1, tm - x[i], i0, i1);
        B0 += i0;
        B1 += i1;
    }
    B1 /= (b - a);
}

// Chebyshev coefficients of exp(-i rho x) on [-1, 1]:
//   exp(-i rho x) = sum_j a_j T_j(x),  a_j = (2 - delta_j0) (-i)^j J_j(rho)
// J_j by Miller's backward recurrence.  Returns m such that terms above m are
// below tol.
inline std::vector<cplx> chebyshev_exp_coeffs(double rho, double tol) {
    rho = std::fabs(rho);
    int mmax = (int)(rho + 30.0 + 12.0 * std::cbrt(rho + 1.0));
    int start = 2 * (mmax / 2) + 40;  // even, comfortably above
    std::vector<double> J(start + 2, 0.0);
    if (rho < 1e-300) {
        J[0] = 1.0;
    } else {
        double jp1 = 0.0, jc = 1e-300;
        std::vector<double> tmp(start + 2, 0.0);
        tmp[start] = jc;
        for (int k = start; k >= 1; --k) {
            double jm1 = (2.0 * k / rho) * jc - jp1;
            jp1 = jc;
            jc = jm1;
            tmp[k - 1] = jc;
            if (std::fabs(jc) > 1e250) {  // rescale to avoid overflow
                for (int q = k - 1; q <= start; ++q) tmp[q] *= 1e-250;
                jc *= 1e-250; jp1 *= 1e-250;
            }
        }
        double norm = tmp[0];
        for (int k = 2; k <= start; k += 2) norm += 2.0 * tmp[k];
        for (int k = 0; k <= start; ++k) J[k] = tmp[k] / norm;
    }
    int m = 0;
    for (int k = mmax; k >= 0; --k) {
        if (std::fabs(J[k]) >= 0.5 * tol) { m = k; break; }
    }
    m = std::max(m, 1);
    std::vector<cplx> a(m + 1);
    const cplx mi[4] = {cplx(1, 0), cplx(0, -1), cplx(-1, 0), cplx(0, 1)};
    for (int j = 0; j <= m; ++j) a[j] = (j == 0 ? 1.0 : 2.0) * mi[j & 3] * J[j];
    return a;
}

}  // namespace pb200

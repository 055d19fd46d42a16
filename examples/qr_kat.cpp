// C++ host-mirror smoke program: the reference's own QR / Cholesky / eigh / svd known-answer tests (src/qr.rs:257-276,
// src/cholesky.rs:209-215, src/eigh.rs:374-381, src/svd.rs:537-543) through include/linfa_b200.hpp.  Build:
//   g++ -std=c++17 -Iinclude examples/qr_kat.cpp -Llinfa_linalg_b200/lib -llinfa_b200 -Wl,-rpath,$PWD/linfa_linalg_b200/lib -o examples/qr_kat
#include <algorithm>
#include <cmath>
#include <cstdio>

#include "linfa_b200.hpp"

using namespace linfa_b200;

int main() {
    Engine eng(0);
    Matrix<double> a(3, 2);
    const double v[6] = {3.2, 1.3, 4.4, 5.2, 1.3, 6.7};
    for (int i = 0; i < 6; ++i) a.data[i] = v[i];
    auto dec = qr_into(eng, a.view());
    auto q = dec.generate_q();
    auto r = dec.into_r();
    const double qe[6] = {0.5720674, -0.4115578, 0.7865927, 0.0301901, 0.2324024, 0.9108835};
    double err = 0;
    for (int i = 0; i < 6; ++i) err = std::fmax(err, std::fabs(q.data[i] - qe[i]));
    err = std::fmax(err, std::fabs(r(0, 0) - 5.594) - 1e-3 + 1e-5);
    Matrix<double> s(3, 3);
    const double sv[9] = {25, 15, -5, 15, 18, 0, -5, 0, 11};
    for (int i = 0; i < 9; ++i) s.data[i] = sv[i];
    cholesky_inplace(eng, s.view());
    const double le[9] = {5, 0, 0, 3, 3, 0, -1, 1, 3};
    for (int i = 0; i < 9; ++i) err = std::fmax(err, std::fabs(s.data[i] - le[i]));
    // src/eigh.rs:374-381 sym_eigvecs1 and src/svd.rs:537-543 svd_test through the whole-driver entry points
    Matrix<double> m(3, 3, 1.0);
    for (int i = 0; i < 3; ++i) m(i, i) = 3.0;
    auto ev = eigh_into(eng, m.view());
    std::sort(ev.first.begin(), ev.first.end());
    const double ee[3] = {2, 2, 5};
    for (int i = 0; i < 3; ++i) err = std::fmax(err, std::fabs(ev.first[i] - ee[i]));
    Matrix<double> d2(2, 2);
    d2(0, 0) = 3.0; d2(1, 1) = -2.0;
    auto sres = svd_into(eng, d2.view(), true, true);
    err = std::fmax(err, std::fabs(std::fmax(sres.sigma[0], sres.sigma[1]) - 3.0));
    err = std::fmax(err, std::fabs(std::fmin(sres.sigma[0], sres.sigma[1]) - 2.0));
    std::printf("max deviation from the reference KATs: %.3e\n", err);
    bool threw = false;
    try { Matrix<double> w(2, 3); qr_into(eng, w.view()); } catch (const NotThin &) { threw = true; }
    return (err < 1e-5 && threw) ? 0 : 1;
}

// Stage schedule of the explicit integrators (host only; shared by api.cu and the host emulation of the kernels, tests/emul).
#pragma once
#include <vector>

#include "mlb_internal.h"

namespace mlb {

// ---------------------------------------------------------------------------------------------------------------
// Stage schedule (numerics/time_integrator.cpp:57-163).  Buffers: U[cur] = solution, the two others are temporaries.
// ---------------------------------------------------------------------------------------------------------------
struct StagePlan { int in, out, base, mode, n_prev, last; double coef, c0, c1, cprev[3]; int kprev[3]; int kstore; };

inline std::vector<StagePlan> make_stage_plan(int cur, int integrator) {
    const int A = cur, B = (cur + 1) % 3, C = (cur + 2) % 3;
    std::vector<StagePlan> p;
    auto mk = [&](int in, int out, int mode, double coef, int kstore) {
        StagePlan s{}; s.in = in; s.out = out; s.base = A; s.mode = mode; s.coef = coef; s.kstore = kstore; return s; };
    if (integrator == MLB_INTEGRATOR_FE) {
        StagePlan s = mk(A, B, 0, 1.0, 0); s.last = 1; p.push_back(s);
    } else if (integrator == MLB_INTEGRATOR_RK4) {
        p.push_back(mk(A, B, 0, 0.5, 0));
        p.push_back(mk(B, C, 0, 0.5, 1));
        p.push_back(mk(C, B, 0, 1.0, 2));
        StagePlan s = mk(B, A, 2, 1.0 / 6.0, 3);
        s.n_prev = 3; s.kprev[0] = 0; s.kprev[1] = 1; s.kprev[2] = 2;
        s.cprev[0] = 1.0 / 6.0; s.cprev[1] = 1.0 / 3.0; s.cprev[2] = 1.0 / 3.0; s.last = 1;
        p.push_back(s);
    } else {
        p.push_back(mk(A, B, 0, 1.0, 0));
        StagePlan s1 = mk(B, C, 1, 0.25, 1); s1.c0 = 3.0 / 4.0; s1.c1 = 1.0 / 4.0; p.push_back(s1);
        StagePlan s2 = mk(C, A, 2, 2.0 / 3.0, 2);
        s2.n_prev = 2; s2.kprev[0] = 0; s2.kprev[1] = 1; s2.cprev[0] = 1.0 / 6.0; s2.cprev[1] = 1.0 / 6.0; s2.last = 1;
        p.push_back(s2);
    }
    return p;
}

}  // namespace mlb

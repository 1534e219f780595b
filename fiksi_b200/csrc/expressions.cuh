// Device-side residual + gradient evaluators for the 11 expression kinds of Fiksi and the two pose rows of ClusteredSystem
// (fiksi/src/constraints/expressions.rs:291-874; formulas tabulated in SURVEY.md App. B).
//
// Arithmetic contract: FP64, no fused multiply-add (Rust never contracts a*b+c; this translation
// unit is compiled with -fmad=false), IEEE division and square root (nvcc defaults
// -prec-div=true -prec-sqrt=true), multiplication order left to right as written in the
// reference.  With that every kind except the two atan2 kinds is bit-identical to the reference
// formulas; CUDA's atan2 is within 2 ulp of glibc's.
#pragma once
#include <cstdint>

namespace fk {
namespace dev {

#define FK_PI 3.14159265358979323846264338327950288

// Point-point distance core, expressions.rs:327-352.
__device__ __forceinline__ double ppd(double p1x, double p1y, double p2x, double p2y, double dist,
                                      double& g0, double& g1, double& g2, double& g3) {
    const double dx = p1x - p2x, dy = p1y - p2y;
    const double d = sqrt(dx * dx + dy * dy);
    const double inv = 1.0 / d;
    g0 = dx * inv;
    g1 = dy * inv;
    g2 = -dx * inv;
    g3 = -dy * inv;
    return d - dist;
}

__device__ __forceinline__ double wrap_pi(double a) {
    if (a > FK_PI) return a - 2.0 * FK_PI;
    if (a < -FK_PI) return a + 2.0 * FK_PI;
    return a;
}

// v: slot values in `variable_indices` order; g: one gradient entry per slot; returns residual.
__device__ __forceinline__ double eval_expression(int kind, const double (&v)[8], double param, double (&g)[8]) {
    double r = 0.0;
    switch (kind) {
        case 0: {  // VariableVariableEquality :291-301
            g[0] = -1.0;
            g[1] = 1.0;
            r = v[1] - v[0];
        } break;
        case 1: {  // PointPointDistance :319-353
            r = ppd(v[0], v[1], v[2], v[3], param, g[0], g[1], g[2], g[3]);
        } break;
        case 2: {  // PointPointPointAngle :372-425
            const double ux = v[0] - v[2], uy = v[1] - v[3];
            const double wx = v[4] - v[2], wy = v[5] - v[3];
            r = wrap_pi(atan2(wy, wx) - atan2(uy, ux)) - param;
            const double ur = 1.0 / (ux * ux + uy * uy);
            const double wr = 1.0 / (wx * wx + wy * wy);
            const double a1x = uy * ur, a1y = -ux * ur;
            const double a3x = -wy * wr, a3y = wx * wr;
            g[0] = a1x; g[1] = a1y;
            g[2] = -a1x - a3x; g[3] = -a1y - a3y;
            g[4] = a3x; g[5] = a3y;
        } break;
        case 3: {  // PointLineIncidence :445-477
            const double ux = v[4] - v[2], uy = v[5] - v[3];
            const double wx = v[0] - v[2], wy = v[1] - v[3];
            r = ux * wy - uy * wx;
            g[0] = -uy; g[1] = ux;
            g[2] = -v[1] + v[5]; g[3] = v[0] - v[4];
            g[4] = wy; g[5] = -wx;
        } break;
        case 4: {  // PointLineDistance :500-544
            const double ux = v[4] - v[2], uy = v[5] - v[3];
            const double wx = v[0] - v[2], wy = v[1] - v[3];
            const double cr = ux * wy - uy * wx;
            const double l2 = ux * ux + uy * uy;
            const double lr = 1.0 / sqrt(l2);
            const double a = cr / l2;
            const double b = -a * ux;
            const double c = v[0] + a * uy;
            r = lr * cr - param;
            g[0] = -lr * uy;
            g[1] = lr * ux;
            g[2] = -lr * (b - v[5] + v[1]);
            g[3] = -lr * (v[4] - c);
            g[4] = lr * (b + wy);
            g[5] = -lr * (c - v[2]);
        } break;
        case 5: {  // PointCircleIncidence :560-576
            r = ppd(v[0], v[1], v[2], v[3], v[4], g[0], g[1], g[2], g[3]);
            g[4] = -1.0;
        } break;
        case 6: {  // SegmentSegmentLengthEquality :593-620
            double a0, a1, a2, a3;
            const double r1 = ppd(v[0], v[1], v[2], v[3], 0.0, a0, a1, a2, a3);
            const double r2 = ppd(v[4], v[5], v[6], v[7], 0.0, g[4], g[5], g[6], g[7]);
            g[0] = -a0; g[1] = -a1; g[2] = -a2; g[3] = -a3;
            r = r2 - r1;
        } break;
        case 7: {  // LineLineAngle :640-696
            const double ux = v[2] - v[0], uy = v[3] - v[1];
            const double wx = v[6] - v[4], wy = v[7] - v[5];
            r = wrap_pi(atan2(wy, wx) - atan2(uy, ux)) - param;
            const double ur = 1.0 / (ux * ux + uy * uy);
            const double wr = 1.0 / (wx * wx + wy * wy);
            const double a = -uy * ur, b = ux * ur, c = wy * wr, d = -wx * wr;
            g[0] = a; g[1] = b; g[2] = -a; g[3] = -b;
            g[4] = c; g[5] = d; g[6] = -c; g[7] = -d;
        } break;
        case 8: {  // LineLineParallelism :713-752
            const double ux = v[2] - v[0], uy = v[3] - v[1];
            const double wx = v[6] - v[4], wy = v[7] - v[5];
            r = wx * uy - wy * ux;
            g[0] = wy; g[1] = -wx; g[2] = -wy; g[3] = wx;
            g[4] = -uy; g[5] = ux; g[6] = uy; g[7] = -ux;
        } break;
        case 9: {  // LineLinePerpendicularity :769-799
            const double ux = v[2] - v[0], uy = v[3] - v[1];
            const double wx = v[6] - v[4], wy = v[7] - v[5];
            r = wx * ux + wy * uy;
            g[0] = -wx; g[1] = -wy; g[2] = wx; g[3] = wy;
            g[4] = -ux; g[5] = -uy; g[6] = ux; g[7] = uy;
        } break;
        case 10: {  // LineCircleTangency :816-874
            const double l1x = v[0], l1y = v[1], l2x = v[2], l2y = v[3], cx = v[4], cy = v[5];
            const double ex = l1x - l2x, ey = l1y - l2y;
            const double len2 = ex * ex + ey * ey;
            const double len = sqrt(len2);
            if (len == 0.0) {  // degenerate line: residual and gradient are all zero (:838-840)
                g[0] = g[1] = g[2] = g[3] = g[4] = g[5] = g[6] = 0.0;
                r = 0.0;
            } else {
                const double lr = 1.0 / len;
                const double area = l1x * (l2y - cy) + l2x * (cy - l1y) + cx * (l1y - l2y);
                r = lr * fabs(area) - v[6];
                // f64::signum: +-1 by sign bit, NaN stays NaN
                const double sg = (area != area) ? area : copysign(1.0, area);
                const double l3r = 1.0 / (len2 * len);
                g[0] = sg * l3r * (len2 * (l2y - cy) + area * (l2x - l1x));
                g[1] = sg * l3r * (len2 * (-l2x + cx) + area * (l2y - l1y));
                g[2] = sg * l3r * (len2 * (cy - l1y) - area * (l2x - l1x));
                g[3] = sg * l3r * (len2 * (l1x - cx) - area * (l2y - l1y));
                g[4] = sg * lr * (l1y - l2y);
                g[5] = sg * lr * (-l1x + l2x);
                g[6] = -1.0;
            }
        } break;
        case 11: case 12: {  // ClusteredSystem's coincidence rows, assemble/mod.rs:538-585 with Pose2D (expressions.rs:1122-1158)
            // slots: rotation, tx, ty, updated coordinate, u, v (the point before the step; fixed, its gradient is unused)
            double s, c;
            sincos(v[0], &s, &c);
            const double u = v[4], w = v[5];
            const double uc = u * c, us = u * s, vc = w * c, vs = w * s;
            if (kind == 11) {
                r = (v[1] + uc - vs) - v[3];
                g[0] = (-us - vc) * 1.0 + (uc - vs) * 0.0;
                g[1] = 1.0;
                g[2] = 0.0;
            } else {
                r = (v[2] + us + vc) - v[3];
                g[0] = (-us - vc) * 0.0 + (uc - vs) * 1.0;
                g[1] = 0.0;
                g[2] = 1.0;
            }
            g[3] = -1.0;
            g[4] = 0.0;
            g[5] = 0.0;
        } break;
        default: break;
    }
    return r;
}

__device__ __forceinline__ int arity_of(int kind) {
    // 2 4 6 6 6 5 8 8 8 8 7 6 6 packed as nibbles
    return (int)((0x6678888566642ull >> (4 * kind)) & 0xF);
}

}  // namespace dev
}  // namespace fk

// Nonlinear models of the robust-MPC hot path and their RK4 discretisation, written once for any scalar type
// (double for roll-outs, Dual for forward-mode derivatives).
//
// Reference behaviour reproduced (file:line in the reference repository):
//   dyn/model.py:15-34      Model.ddyn  -- classical RK4, default step h = 0.05
//   dyn/pendulum.py:26-44   cart-pole ODE
//   dyn/quadrotor.py:108-174 quadrotor ODE (quaternion [w,x,y,z])
//   dyn/rocket.py:165-254   rockETH ODE incl. gimbal map 2*atan((v - sqrt(w^2+v^2-u^2))/(u+w))
#pragma once
#ifdef __CUDACC__
#include <cuda_runtime.h>
#else   // plain C++ build (oracle/cpu_port): the ODEs are ordinary host functions
#ifndef __host__
#define __host__
#define __device__
#endif
#endif
#include <math.h>

namespace rnmpc {

// ---- forward-mode dual number (value + one tangent) ----------------------------------------------------------
struct Dual {
    double v, d;
    __host__ __device__ Dual() : v(0.0), d(0.0) {}
    __host__ __device__ Dual(double v_) : v(v_), d(0.0) {}
    __host__ __device__ Dual(double v_, double d_) : v(v_), d(d_) {}
};
__host__ __device__ inline Dual operator+(Dual a, Dual b) { return Dual(a.v + b.v, a.d + b.d); }
__host__ __device__ inline Dual operator-(Dual a, Dual b) { return Dual(a.v - b.v, a.d - b.d); }
__host__ __device__ inline Dual operator-(Dual a) { return Dual(-a.v, -a.d); }
__host__ __device__ inline Dual operator*(Dual a, Dual b) { return Dual(a.v * b.v, fma(a.v, b.d, a.d * b.v)); }
// Quotients: ONE division (the reciprocal of the divisor's value) per dual quotient; a division by a plain number -- in the ODEs
// below always a compile-time constant -- is a multiplication by its reciprocal.  (Measured on linearize_kernel: IEEE divisions were
// a third of its instructions.)  Values differ from the plain-double path by an ulp at most.
__host__ __device__ inline Dual operator/(Dual a, Dual b) {
    const double r = 1.0 / b.v, q = a.v * r;
    return Dual(q, (a.d - q * b.d) * r);
}
__host__ __device__ inline Dual operator+(Dual a, double b) { return Dual(a.v + b, a.d); }
__host__ __device__ inline Dual operator+(double a, Dual b) { return Dual(a + b.v, b.d); }
__host__ __device__ inline Dual operator-(Dual a, double b) { return Dual(a.v - b, a.d); }
__host__ __device__ inline Dual operator-(double a, Dual b) { return Dual(a - b.v, -b.d); }
__host__ __device__ inline Dual operator*(Dual a, double b) { return Dual(a.v * b, a.d * b); }
__host__ __device__ inline Dual operator*(double a, Dual b) { return Dual(a * b.v, a * b.d); }
__host__ __device__ inline Dual operator/(Dual a, double b) { const double r = 1.0 / b; return Dual(a.v * r, a.d * r); }
__host__ __device__ inline Dual operator/(double a, Dual b) {
    const double r = 1.0 / b.v, q = a * r;
    return Dual(q, -q * b.d * r);
}
__host__ __device__ inline Dual m_sin(Dual a) { double s, c; sincos(a.v, &s, &c); return Dual(s, c * a.d); }
__host__ __device__ inline Dual m_cos(Dual a) { double s, c; sincos(a.v, &s, &c); return Dual(c, -s * a.d); }
__host__ __device__ inline void m_sincos(Dual a, Dual& s_, Dual& c_) {
    double s, c; sincos(a.v, &s, &c);
    s_ = Dual(s, c * a.d); c_ = Dual(c, -s * a.d);
}
__host__ __device__ inline Dual m_sqrt(Dual a) {
#ifdef __CUDA_ARCH__
    // reciprocal square root + one correction step for the value (no division, no IEEE square root)
    const double rs = rsqrt(a.v);
    double r = a.v * rs;
    r = fma(0.5 * rs, fma(-r, r, a.v), r);
    return Dual(r, 0.5 * a.d * rs);
#else
    const double r = sqrt(a.v);
    return Dual(r, 0.5 * a.d / r);
#endif
}
__host__ __device__ inline Dual m_atan(Dual a) { return Dual(atan(a.v), a.d / (1.0 + a.v * a.v)); }
__host__ __device__ inline double m_sin(double a) { return sin(a); }
__host__ __device__ inline double m_cos(double a) { return cos(a); }
__host__ __device__ inline void m_sincos(double a, double& s, double& c) { sincos(a, &s, &c); }
__host__ __device__ inline double m_sqrt(double a) { return sqrt(a); }
__host__ __device__ inline double m_atan(double a) { return atan(a); }
__host__ __device__ inline double value_of(double a) { return a; }
__host__ __device__ inline double value_of(Dual a) { return a.v; }
__host__ __device__ inline double tangent_of(double) { return 0.0; }
__host__ __device__ inline double tangent_of(Dual a) { return a.d; }

enum ModelId { MODEL_PENDULUM = 0, MODEL_QUADROTOR = 1, MODEL_ROCKET = 2 };

// ---- shared rigid-body pieces ----------------------------------------------------------------------------------
template <class T>
__host__ __device__ inline void quat_kinematics(T qw, T qx, T qy, T qz, T wx, T wy, T wz, T* qd) {
    // dyn/rocket.py:214-223 : q_dot = 0.5 * Omega(omega) * [w,x,y,z]
    qd[0] = 0.5 * (-(wx * qx) - wy * qy - wz * qz);
    qd[1] = 0.5 * (wx * qw + wz * qy - wy * qz);
    qd[2] = 0.5 * (wy * qw - wz * qx + wx * qz);
    qd[3] = 0.5 * (wz * qw + wy * qx - wx * qy);
}

// ---- pendulum (dyn/pendulum.py:26-44) ----------------------------------------------------------------------------
// Kinematic-chain states: the first NPAS states do not enter the right-hand side except that state t (CHAIN <= t < NPAS) is the
// derivative of state t - CHAIN.  Their columns of A_k = d rk4 / dx are known in closed form (e_t, or e_t + h e_{t-CHAIN}: the RK4
// weights sum to one) and the linearisation kernel writes them without differentiating anything.
struct PendulumModel {
    static constexpr int NX = 4, NU = 1, ID = MODEL_PENDULUM;
    static constexpr int NPAS = 2, CHAIN = 1;          // cart position; cart velocity -> position
    template <class T>
    __host__ __device__ static inline void ode(const T* X, const T* u, T* xd) {
        const double m1 = 1.0, m2 = 0.1, l = 0.5, g = 9.81;
        T s, c;
        m_sincos(X[2], s, c);
        T thd = X[3];
        T den = m1 + m2 * (1.0 - c * c);
        xd[0] = X[1];
        xd[1] = (u[0] + m2 * l * (thd * thd) * s - m2 * g * s * c) / den;
        xd[2] = thd;
        xd[3] = (-(u[0] * c) - m2 * l * (thd * thd) * s * c + (m1 + m2) * g * s) / (l * den);
    }
};

// ---- quadrotor (dyn/quadrotor.py:108-174) ---------------------------------------------------------------------
struct QuadrotorModel {
    static constexpr int NX = 13, NU = 4, ID = MODEL_QUADROTOR;
    static constexpr int NPAS = 6, CHAIN = 3;          // position; velocity -> position
    template <class T>
    __host__ __device__ static inline void ode(const T* X, const T* u, T* xd) {
        const double m = 1.0, g = 9.81, l = 0.15, Jx = 0.02, Jy = 0.02, Jz = 0.04, kM = 0.01;
        T qw = X[6], qx = X[7], qy = X[8], qz = X[9];
        T wx = X[10], wy = X[11], wz = X[12];
        T Fz = u[0] + u[1] + u[2] + u[3];
        T r02 = 2.0 * qx * qz + 2.0 * qy * qw;
        T r12 = 2.0 * qy * qz - 2.0 * qx * qw;
        T r22 = 1.0 - 2.0 * (qx * qx) - 2.0 * (qy * qy);
        xd[0] = X[3]; xd[1] = X[4]; xd[2] = X[5];
        xd[3] = (1.0 / m) * (r02 * Fz);
        xd[4] = (1.0 / m) * (r12 * Fz);
        xd[5] = (1.0 / m) * (r22 * Fz) - g;
        quat_kinematics(qw, qx, qy, qz, wx, wy, wz, xd + 6);
        T tx = l * (u[1] - u[3]);
        T ty = l * (u[2] - u[0]);
        T tz = kM * (u[0] - u[1] + u[2] - u[3]);
        xd[10] = (tx - (wy * (Jz * wz) - wz * (Jy * wy))) / Jx;
        xd[11] = (ty - (wz * (Jx * wx) - wx * (Jz * wz))) / Jy;
        xd[12] = (tz - (wx * (Jy * wy) - wy * (Jx * wx))) / Jz;
    }
};

// ---- rockETH rocket (dyn/rocket.py:165-254) ------------------------------------------------------------------------
struct RocketModel {
    static constexpr int NX = 17, NU = 4, ID = MODEL_ROCKET;
    static constexpr int NPAS = 6, CHAIN = 3;          // position; velocity -> position
    template <class T>
    __host__ __device__ static inline T gimbal(T sn, T cs, T cos_tilt) {
        // dyn/rocket.py:246-254, with sin/cos of the servo angle passed in
        const double a = 5.0, b = 35.2, c = 33.0, d = 28.0, e = 35.2;
        T iv1 = d + a * cs;
        T iv2 = e - a * sn;
        T uu = (b * b - c * c) - iv1 * iv1 - iv2 * iv2;
        T vv = 2.0 * c * cos_tilt * iv2;
        T ww = -2.0 * c * iv1;
        T iv3 = ww * ww + vv * vv - uu * uu;
        return 2.0 * m_atan((vv - m_sqrt(iv3)) / (uu + ww));
    }
    template <class T>
    __host__ __device__ static inline void ode(const T* X, const T* u, T* xd) {
        const double mass = 1.16, grav = 9.81, Jxx = 0.0021, Jyy = 0.1, Jzz = 0.1, off = 0.42;
        const double tauT = 0.06, tauS = 0.10, hover = 11.3796;
        T qw = X[6], qx = X[7], qy = X[8], qz = X[9];
        T wx = X[10], wy = X[11], wz = X[12];
        T Tm = X[13] + hover;
        T s1, c1, s2, c2;
        m_sincos(X[15], s1, c1);
        m_sincos(X[16], s2, c2);
        T g1 = gimbal(s1, c1, T(1.0));
        T sg1, cg1;
        m_sincos(g1, sg1, cg1);
        T g2 = gimbal(s2, c2, cg1);
        T sg2, cg2;
        m_sincos(g2, sg2, cg2);
        T Fx = -(Tm * sg1 * cg2);
        T Fy = Tm * sg2;
        T Fz = Tm * cg1 * cg2;
        T r00 = 1.0 - 2.0 * (qy * qy) - 2.0 * (qz * qz);
        T r01 = 2.0 * qx * qy - 2.0 * qz * qw;
        T r02 = 2.0 * qx * qz + 2.0 * qy * qw;
        T r10 = 2.0 * qx * qy + 2.0 * qz * qw;
        T r11 = 1.0 - 2.0 * (qx * qx) - 2.0 * (qz * qz);
        T r12 = 2.0 * qy * qz - 2.0 * qx * qw;
        T r20 = 2.0 * qx * qz - 2.0 * qy * qw;
        T r21 = 2.0 * qy * qz + 2.0 * qx * qw;
        T r22 = 1.0 - 2.0 * (qx * qx) - 2.0 * (qy * qy);
        xd[0] = X[3]; xd[1] = X[4]; xd[2] = X[5];
        xd[3] = (1.0 / mass) * (r00 * Fx + r01 * Fy + r02 * Fz);
        xd[4] = (1.0 / mass) * (r10 * Fx + r11 * Fy + r12 * Fz);
        xd[5] = (1.0 / mass) * (r20 * Fx + r21 * Fy + r22 * Fz) - grav;
        quat_kinematics(qw, qx, qy, qz, wx, wy, wz, xd + 6);
        // torque = [0,0,-off] x F = (off*Fy, -off*Fx, 0);  J wdot = torque - w x (J w)
        xd[10] = (off * Fy - (wy * (Jzz * wz) - wz * (Jyy * wy))) / Jxx;
        xd[11] = (-(off * Fx) - (wz * (Jxx * wx) - wx * (Jzz * wz))) / Jyy;
        xd[12] = (-(wx * (Jyy * wy) - wy * (Jxx * wx))) / Jzz;
        xd[13] = (u[0] + hover - Tm) / tauT;
        xd[14] = (u[1] - X[14]) / tauT;
        xd[15] = (u[2] - X[15]) / tauS;
        xd[16] = (u[3] - X[16]) / tauS;
    }
};

// ---- RK4 step (dyn/model.py:26-32) ------------------------------------------------------------------------------
template <class M, class T>
__host__ __device__ inline void rk4_step(const T* x, const T* u, double h, T* xp) {
    constexpr int NX = M::NX;
    T k1[NX], k2[NX], k3[NX], k4[NX], xt[NX];
    M::ode(x, u, k1);
#pragma unroll
    for (int i = 0; i < NX; ++i) xt[i] = x[i] + (0.5 * h) * k1[i];
    M::ode(xt, u, k2);
#pragma unroll
    for (int i = 0; i < NX; ++i) xt[i] = x[i] + (0.5 * h) * k2[i];
    M::ode(xt, u, k3);
#pragma unroll
    for (int i = 0; i < NX; ++i) xt[i] = x[i] + h * k3[i];
    M::ode(xt, u, k4);
#pragma unroll
    for (int i = 0; i < NX; ++i) xp[i] = x[i] + (1.0 / 6.0) * (k1[i] + 2.0 * k2[i] + 2.0 * k3[i] + k4[i]) * h;
}

}  // namespace rnmpc

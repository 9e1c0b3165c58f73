// Host restatement of the device's expf_libm (csrc/fwgpu_kernels.cuh): glibc's expf algorithm.
// tests/test_expf.py compares it bit for bit with the C library's expf, which is what the reference's
// logistic() calls through Rust's f32::exp (block_loss_functions.rs:15-17).
#include "../../../include/fwhost.h"
#include <cstdint>
#include <cstring>

static const uint64_t kTab[32] = {
    0x3ff0000000000000ull, 0x3fefd9b0d3158574ull, 0x3fefb5586cf9890full, 0x3fef9301d0125b51ull,
    0x3fef72b83c7d517bull, 0x3fef54873168b9aaull, 0x3fef387a6e756238ull, 0x3fef1e9df51fdee1ull,
    0x3fef06fe0a31b715ull, 0x3feef1a7373aa9cbull, 0x3feedea64c123422ull, 0x3feece086061892dull,
    0x3feebfdad5362a27ull, 0x3feeb42b569d4f82ull, 0x3feeab07dd485429ull, 0x3feea47eb03a5585ull,
    0x3feea09e667f3bcdull, 0x3fee9f75e8ec5f74ull, 0x3feea11473eb0187ull, 0x3feea589994cce13ull,
    0x3feeace5422aa0dbull, 0x3feeb737b0cdc5e5ull, 0x3feec49182a3f090ull, 0x3feed503b23e255dull,
    0x3feee89f995ad3adull, 0x3feeff76f2fb5e47ull, 0x3fef199bdd85529cull, 0x3fef3720dcef9069ull,
    0x3fef5818dcfba487ull, 0x3fef7c97337b9b5full, 0x3fefa4afa2a490daull, 0x3fefd0765b6e4540ull};

extern "C" float fwhost_expf_libm(float x)
{
    const double InvLn2N = 0x1.71547652b82fep+0 * 32.0, SHIFT = 0x1.8p+52;
    const double C0 = 0x1.c6af84b912394p-5 / 32.0 / 32.0 / 32.0, C1 = 0x1.ebfce50fac4f3p-3 / 32.0 / 32.0, C2 = 0x1.62e42ff0c52d6p-1 / 32.0;
    volatile double z = InvLn2N * (double)x;
    volatile double kd = z + SHIFT;
    uint64_t ki;
    double kdv = kd;
    std::memcpy(&ki, &kdv, 8);
    volatile double kd2 = kd - SHIFT;
    volatile double r = z - kd2;
    uint64_t t = kTab[ki & 31] + (ki << 47);
    double sc;
    std::memcpy(&sc, &t, 8);
    volatile double m0 = C0 * r;
    volatile double zz = m0 + C1;
    volatile double r2 = r * r;
    volatile double m1 = C2 * r;
    volatile double y = m1 + 1.0;
    volatile double m2 = zz * r2;
    volatile double y2 = m2 + y;
    volatile double y3 = y2 * sc;
    return (float)y3;
}

extern "C" void fwhost_expf_libm_array(const float *in, float *out, uint64_t n)
{
    for (uint64_t i = 0; i < n; i++) out[i] = fwhost_expf_libm(in[i]);
}

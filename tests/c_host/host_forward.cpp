// A host with no Python and no torch in the process: CUDA runtime + the C ABI only (include/omg_b200.h).
//   1. a Linear with bias and residual through omg_gemm on cudaMalloc'ed buffers, checked against a CPU loop;
//   2. the same call recorded into a launch plan (omg_plan_*) together with an omg_layernorm, replayed on new input data and
//      checked again - the "forward as a handle" protocol of INTEGRATION.md B2.
// Built and run by tests/test_c_host_gpu.py:  g++ host_forward.cpp -I include -I $CUDA/include -L omg_b200/lib -lomg_b200 -lcudart
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "omg_b200.h"

#define CK(x)                                                                  \
    do {                                                                       \
        cudaError_t e_ = (x);                                                  \
        if (e_ != cudaSuccess) {                                               \
            fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_));           \
            return 2;                                                          \
        }                                                                      \
    } while (0)
#define OMG(x)                                                   \
    do {                                                         \
        if ((x) != 0) {                                          \
            fprintf(stderr, "%s: %s\n", #x, omg_last_error());   \
            return 3;                                            \
        }                                                        \
    } while (0)

static float frand(unsigned* s) {
    *s = *s * 1664525u + 1013904223u;
    return ((*s >> 8) & 0xffff) / 32768.0f - 1.0f;
}

static double rel_l2(const std::vector<float>& a, const std::vector<float>& b) {
    double n = 0, d = 0;
    for (size_t i = 0; i < a.size(); ++i) {
        d += (double)(a[i] - b[i]) * (a[i] - b[i]);
        n += (double)b[i] * b[i];
    }
    return sqrt(d / n);
}

int main() {
    const int M = 512, N = 320, K = 256;
    unsigned seed = 7;
    std::vector<__half> hx(M * K), hw(N * K), hb(N), hr(M * N);
    auto fill = [&](std::vector<__half>& v, float s) {
        for (auto& e : v) e = __float2half(frand(&seed) * s);
    };
    fill(hx, 1.0f);
    fill(hw, 0.0625f);
    fill(hb, 0.5f);
    fill(hr, 1.0f);
    __half *dx, *dw, *db, *dr, *dy, *dz, *dg, *dbeta;
    CK(cudaMalloc(&dx, hx.size() * 2));
    CK(cudaMalloc(&dw, hw.size() * 2));
    CK(cudaMalloc(&db, hb.size() * 2));
    CK(cudaMalloc(&dr, hr.size() * 2));
    CK(cudaMalloc(&dy, (size_t)M * N * 2));
    CK(cudaMalloc(&dz, (size_t)M * N * 2));
    CK(cudaMalloc(&dg, N * 2));
    CK(cudaMalloc(&dbeta, N * 2));
    std::vector<__half> ones(N, __float2half(1.0f)), zeros(N, __float2half(0.0f));
    CK(cudaMemcpy(dg, ones.data(), N * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dbeta, zeros.data(), N * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dw, hw.data(), hw.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dr, hr.data(), hr.size() * 2, cudaMemcpyHostToDevice));
    cudaStream_t stream;
    CK(cudaStreamCreate(&stream));

    omg_gemm_desc d;
    memset(&d, 0, sizeof(d));
    // a [M, K] matrix as a one-image, one-row-high channels-last view
    d.a[0].ptr = dx; d.a[0].C = K; d.a[0].W = M; d.a[0].H = 1; d.a[0].B = 1; d.a[0].sw = K; d.a[0].sh = (int64_t)K * M; d.a[0].sb = (int64_t)K * M;
    d.n_a = 1;
    d.segs[0].a_idx = 0; d.segs[0].k_len = K;
    d.n_segs = 1;
    d.w = dw; d.N = N; d.Ktot = K;
    d.d.ptr = dy; d.d.C = N; d.d.W = M; d.d.H = 1; d.d.B = 1; d.d.sw = N; d.d.sh = (int64_t)N * M; d.d.sb = (int64_t)N * M;
    d.bias = db;
    d.residual = dr; d.residual_ld = N;
    d.epilogue = OMG_EPI_NONE;

    auto reference = [&](std::vector<float>& y, std::vector<float>& z) {
        y.assign((size_t)M * N, 0.f);
        z.assign((size_t)M * N, 0.f);
        for (int m = 0; m < M; ++m) {
            double s = 0, q = 0;
            for (int n = 0; n < N; ++n) {
                float acc = 0.f;
                for (int k = 0; k < K; ++k) acc += __half2float(hx[(size_t)m * K + k]) * __half2float(hw[(size_t)n * K + k]);
                const float v = __half2float(__float2half(acc + __half2float(hb[n]) + __half2float(hr[(size_t)m * N + n])));
                y[(size_t)m * N + n] = v;
                s += v;
                q += (double)v * v;
            }
            const double mu = s / N, rstd = 1.0 / sqrt(q / N - mu * mu + 1e-5);
            for (int n = 0; n < N; ++n) z[(size_t)m * N + n] = (float)((y[(size_t)m * N + n] - mu) * rstd);
        }
    };
    auto fetch = [&](const __half* dev, std::vector<float>& out) {
        std::vector<__half> h((size_t)M * N);
        cudaMemcpy(h.data(), dev, h.size() * 2, cudaMemcpyDeviceToHost);
        out.resize(h.size());
        for (size_t i = 0; i < h.size(); ++i) out[i] = __half2float(h[i]);
    };

    // 1. direct launches, recorded into a plan at the same time
    omg_plan* plan = omg_plan_create();
    if (!plan) return 4;
    CK(cudaMemcpy(dx, hx.data(), hx.size() * 2, cudaMemcpyHostToDevice));
    OMG(omg_plan_record_begin(plan));
    OMG(omg_gemm(&d, stream));
    OMG(omg_layernorm(dy, dg, dbeta, dz, M, N, 1e-5f, stream));
    OMG(omg_plan_record_end(plan));
    CK(cudaStreamSynchronize(stream));
    std::vector<float> ry, rz, gy, gz;
    reference(ry, rz);
    fetch(dy, gy);
    fetch(dz, gz);
    const double e1 = rel_l2(gy, ry), e2 = rel_l2(gz, rz);
    printf("direct: gemm rel_l2 %.3e layernorm rel_l2 %.3e plan length %d launches %llu\n", e1, e2, omg_plan_length(plan),
           (unsigned long long)omg_launch_count());
    if (!(e1 < 2e-3 && e2 < 2e-3) || omg_plan_length(plan) != 2) return 5;

    // 2. new input, outputs cleared, the descriptor on the stack scribbled over: replay from the handle
    fill(hx, 1.0f);
    memset(&d, 0xff, sizeof(d));
    CK(cudaMemcpy(dx, hx.data(), hx.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemset(dy, 0, (size_t)M * N * 2));
    CK(cudaMemset(dz, 0, (size_t)M * N * 2));
    OMG(omg_plan_run(plan, stream));
    CK(cudaStreamSynchronize(stream));
    reference(ry, rz);
    fetch(dy, gy);
    fetch(dz, gz);
    const double e3 = rel_l2(gy, ry), e4 = rel_l2(gz, rz);
    printf("replay: gemm rel_l2 %.3e layernorm rel_l2 %.3e launches %llu\n", e3, e4, (unsigned long long)omg_launch_count());
    if (!(e3 < 2e-3 && e4 < 2e-3) || omg_launch_count() != 4) return 6;
    omg_plan_destroy(plan);
    printf("C host OK (%s)\n", omg_version());
    return 0;
}

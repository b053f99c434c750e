// Per-Gaussian backward (sm_100a): one streaming kernel that turns the 9-float accumulator of the tile pass
// into the FINAL dense gradient tensors.
//
// Replaces, fused:  computeCov2DCUDA   RAST/cuda_rasterizer/backward.cu:144-274
//                   preprocessCUDA bwd backward.cu:346-396 (+ SH backward :20-139, cov3D backward :278-341)
//                   the nine torch::zeros fills of the binding (RAST/rasterize_points.cu:154-162)
// Every output row is written exactly once (zeros for invisible Gaussians), so the caller passes
// uninitialised memory; dL_dconic / dL_dcov3D / dL_dcolor never exist as dense [P,*] tensors unless asked for.
// HBM-bound: ~248 B written + 4 B read per Gaussian (+ ~290 B read per visible one).  All stores are coalesced:
// per-thread results are staged in shared memory and written by the warp as contiguous 128-bit rows
// (a warp's 32 SH-gradient rows are one contiguous 6 KB block of dL_dsh).
#include "gs_common.cuh"

namespace {

constexpr int kT = 256;
constexpr int kWarps = kT / 32;

// write 32 rows x NF floats, staged in s[32*NF] (row-major), to dst[(row0 + r) * NF + c]; coalesced
template <int NF>
__device__ __forceinline__ void warp_store_rows(float* __restrict__ dst, const float* s, long long row0, long long P,
                                                int lane) {
    const long long base = row0 * NF;
    const long long lim = P * NF;
#pragma unroll
    for (int k = 0; k < NF; k++) {
        const long long e = base + lane + 32 * k;
        if (e < lim) dst[e] = s[lane + 32 * k];
    }
}

// Expand the staged (basis[16], dRGB[3]) of `rows` Gaussians into their contiguous dL_dsh block with coalesced
// 128-bit stores.  CM3 > 0: compile-time row length (M*3), so the index arithmetic is mul/shift, not division.
template <int CM3>
__device__ __forceinline__ void store_dsh(float* __restrict__ dst, const float* sb, const float* sr, int rows,
                                          int lane, int rt_m3 = 0) {
    const int M3 = CM3 > 0 ? CM3 : rt_m3;
    const int total = rows * M3;
    if ((M3 & 3) == 0) {
        const int total4 = total >> 2;
        for (int f = lane; f < total4; f += 32) {
            float o[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int e = 4 * f + u;
                const int gg = e / M3, rem = e - gg * M3;
                const int k = rem / 3, ch = rem - 3 * k;
                o[u] = (k < 16 ? sb[gg * 16 + k] : 0.f) * sr[gg * 3 + ch];
            }
            reinterpret_cast<float4*>(dst)[f] = make_float4(o[0], o[1], o[2], o[3]);
        }
    } else {
        for (int e = lane; e < total; e += 32) {
            const int gg = e / M3, rem = e - gg * M3;
            const int k = rem / 3, ch = rem - 3 * k;
            dst[e] = (k < 16 ? sb[gg * 16 + k] : 0.f) * sr[gg * 3 + ch];
        }
    }
}

__global__ void __launch_bounds__(kT)
k_gauss_bwd(const GsView v, const int* __restrict__ radii, const float* __restrict__ means3D,
            const float* __restrict__ shs, const float* __restrict__ scales, const float* __restrict__ rotations,
            const float* __restrict__ cov3D_precomp, const float4* __restrict__ rec, float4* __restrict__ acc,
            const GsGradPtrs g) {
    __shared__ float s_basis[kWarps][32 * 16];
    __shared__ float s_rgb[kWarps][32 * 3];
    __shared__ float s_stage[kWarps][32 * 6];
    __shared__ GsCam cam;
    gs_load_cam(v, &cam);
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const long long row0 = (long long)blockIdx.x * kT + wid * 32;
    const long long i = row0 + lane;
    const int P = v.P;

    bool vis = false;
    if (i < P) vis = radii[i] > 0;

    float dm3x = 0.f, dm3y = 0.f, dm3z = 0.f, dm2x = 0.f, dm2y = 0.f, dop = 0.f;
    float dsx = 0.f, dsy = 0.f, dsz = 0.f;
    float4 drot = make_float4(0.f, 0.f, 0.f, 0.f);
    float dcol0 = 0.f, dcol1 = 0.f, dcol2 = 0.f;
    float dcov[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float bs[16];
#pragma unroll
    for (int k = 0; k < 16; k++) bs[k] = 0.f;
    float dRGB0 = 0.f, dRGB1 = 0.f, dRGB2 = 0.f;

    if (vis) {
        float4* aa = acc + (size_t)3 * i;
        const float4 a0 = aa[0], a1 = aa[1];
        const float a2x = reinterpret_cast<const float*>(aa + 2)[0];
        const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
        aa[0] = z4; aa[1] = z4; aa[2] = z4;                 // re-arm the accumulator for the next backward
        dm2x = a0.x; dm2y = a0.y; dop = a1.y;
        const float dcx = a0.z, dcy = a0.w, dcz = a1.x;     // dL_dconic (a, b, c)
        dcol0 = a1.z; dcol1 = a1.w; dcol2 = a2x;
        const uint32_t clamped = __float_as_uint(__ldg(reinterpret_cast<const float*>(rec + (size_t)3 * i + 2) + 2));

        const float3 p = make_float3(means3D[3 * i], means3D[3 * i + 1], means3D[3 * i + 2]);
        float c6[6];
        float3 sc = make_float3(0.f, 0.f, 0.f);
        float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
        if (cov3D_precomp) {
#pragma unroll
            for (int k = 0; k < 6; k++) c6[k] = cov3D_precomp[6 * i + k];
        } else {
            sc = make_float3(scales[3 * i], scales[3 * i + 1], scales[3 * i + 2]);
            q = __ldg(reinterpret_cast<const float4*>(rotations) + i);
            gs_cov3d(sc, v.scale_modifier, q, c6);
        }
        // ---- computeCov2DCUDA (backward.cu:144-274)
        GsCov2D cc;
        gs_cov2d(p, v, cam.vm, c6, cc);
        const float a = cc.a, b = cc.b, c = cc.c;
        const float denom = a * c - b * b;
        float dL_da = 0.f, dL_db = 0.f, dL_dc = 0.f;
        const float denom2inv = 1.0f / ((denom * denom) + 0.0000001f);
        const float (*Tm)[3] = cc.A;
        if (denom2inv != 0.f) {
            dL_da = denom2inv * (-c * c * dcx + 2 * b * c * dcy + (denom - a * c) * dcz);
            dL_dc = denom2inv * (-a * a * dcz + 2 * a * b * dcy + (denom - a * c) * dcx);
            dL_db = denom2inv * 2 * (b * c * dcx - (denom + 2 * b * b) * dcy + a * b * dcz);
            dcov[0] = (Tm[0][0] * Tm[0][0] * dL_da + Tm[0][0] * Tm[1][0] * dL_db + Tm[1][0] * Tm[1][0] * dL_dc);
            dcov[3] = (Tm[0][1] * Tm[0][1] * dL_da + Tm[0][1] * Tm[1][1] * dL_db + Tm[1][1] * Tm[1][1] * dL_dc);
            dcov[5] = (Tm[0][2] * Tm[0][2] * dL_da + Tm[0][2] * Tm[1][2] * dL_db + Tm[1][2] * Tm[1][2] * dL_dc);
            dcov[1] = 2 * Tm[0][0] * Tm[0][1] * dL_da + (Tm[0][0] * Tm[1][1] + Tm[0][1] * Tm[1][0]) * dL_db + 2 * Tm[1][0] * Tm[1][1] * dL_dc;
            dcov[2] = 2 * Tm[0][0] * Tm[0][2] * dL_da + (Tm[0][0] * Tm[1][2] + Tm[0][2] * Tm[1][0]) * dL_db + 2 * Tm[1][0] * Tm[1][2] * dL_dc;
            dcov[4] = 2 * Tm[0][2] * Tm[0][1] * dL_da + (Tm[0][1] * Tm[1][2] + Tm[0][2] * Tm[1][1]) * dL_db + 2 * Tm[1][1] * Tm[1][2] * dL_dc;
        }
        const float S[3][3] = {{c6[0], c6[1], c6[2]}, {c6[1], c6[3], c6[4]}, {c6[2], c6[4], c6[5]}};
        float dT[2][3];
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const float u0 = Tm[0][0] * S[k][0] + Tm[0][1] * S[k][1] + Tm[0][2] * S[k][2];
            const float u1 = Tm[1][0] * S[k][0] + Tm[1][1] * S[k][1] + Tm[1][2] * S[k][2];
            dT[0][k] = 2 * u0 * dL_da + u1 * dL_db;
            dT[1][k] = 2 * u1 * dL_dc + u0 * dL_db;
        }
        const float* vm = cam.vm;
        const float dJ00 = vm[0] * dT[0][0] + vm[4] * dT[0][1] + vm[8] * dT[0][2];
        const float dJ02 = vm[2] * dT[0][0] + vm[6] * dT[0][1] + vm[10] * dT[0][2];
        const float dJ11 = vm[1] * dT[1][0] + vm[5] * dT[1][1] + vm[9] * dT[1][2];
        const float dJ12 = vm[2] * dT[1][0] + vm[6] * dT[1][1] + vm[10] * dT[1][2];
        const float tz = 1.f / cc.tz, tz2 = tz * tz, tz3 = tz2 * tz;
        const float dtx = cc.xmul * -v.focal_x * tz2 * dJ02;
        const float dty = cc.ymul * -v.focal_y * tz2 * dJ12;
        const float dtz = -v.focal_x * tz2 * dJ00 - v.focal_y * tz2 * dJ11 + (2 * v.focal_x * cc.tx) * tz3 * dJ02 +
                          (2 * v.focal_y * cc.ty) * tz3 * dJ12;
        dm3x = vm[0] * dtx + vm[1] * dty + vm[2] * dtz;     // assignment (backward.cu:273)
        dm3y = vm[4] * dtx + vm[5] * dty + vm[6] * dtz;
        dm3z = vm[8] * dtx + vm[9] * dty + vm[10] * dtz;

        // ---- preprocessCUDA backward (backward.cu:346-396): projection Jacobian
        const float* pj = cam.pm;
        const float4 m_hom = gs_xf4x4(p, pj);
        const float m_w = 1.0f / (m_hom.w + 0.0000001f);
        const float mul1 = (pj[0] * p.x + pj[4] * p.y + pj[8] * p.z + pj[12]) * m_w * m_w;
        const float mul2 = (pj[1] * p.x + pj[5] * p.y + pj[9] * p.z + pj[13]) * m_w * m_w;
        dm3x += (pj[0] * m_w - pj[3] * mul1) * dm2x + (pj[1] * m_w - pj[3] * mul2) * dm2y;
        dm3y += (pj[4] * m_w - pj[7] * mul1) * dm2x + (pj[5] * m_w - pj[7] * mul2) * dm2y;
        dm3z += (pj[8] * m_w - pj[11] * mul1) * dm2x + (pj[9] * m_w - pj[11] * mul2) * dm2y;

        // ---- SH backward (backward.cu:20-139)
        if (shs) {
            const float3 d0 = make_float3(p.x - cam.campos[0], p.y - cam.campos[1], p.z - cam.campos[2]);
            const float len = sqrtf(d0.x * d0.x + d0.y * d0.y + d0.z * d0.z);
            const float x = d0.x / len, y = d0.y / len, z = d0.z / len;
            dRGB0 = dcol0 * ((clamped & 1u) ? 0.f : 1.f);
            dRGB1 = dcol1 * ((clamped & 2u) ? 0.f : 1.f);
            dRGB2 = dcol2 * ((clamped & 4u) ? 0.f : 1.f);
            gs_sh_basis(v.D, x, y, z, bs);
            const float* sh = shs + (size_t)i * v.M * 3;
            float ddx = 0.f, ddy = 0.f, ddz = 0.f;          // dL_ddir
            const int D = v.D;
            if (D > 0) {
                const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
                const float dR[3] = {dRGB0, dRGB1, dRGB2};
                float ax[3], ay[3], az[3];
#pragma unroll
                for (int ch = 0; ch < 3; ch++) {
#define SHK(k) __ldg(sh + 3 * (k) + ch)
                    float vx = -GS_C1 * SHK(3), vy = -GS_C1 * SHK(1), vz = GS_C1 * SHK(2);
                    if (D > 1) {
                        vx += GS_C2_0 * y * SHK(4) + GS_C2_2 * 2.f * -x * SHK(6) + GS_C2_3 * z * SHK(7) + GS_C2_4 * 2.f * x * SHK(8);
                        vy += GS_C2_0 * x * SHK(4) + GS_C2_1 * z * SHK(5) + GS_C2_2 * 2.f * -y * SHK(6) + GS_C2_4 * 2.f * -y * SHK(8);
                        vz += GS_C2_1 * y * SHK(5) + GS_C2_2 * 2.f * 2.f * z * SHK(6) + GS_C2_3 * x * SHK(7);
                        if (D > 2) {
                            vx += (GS_C3_0 * SHK(9) * 3.f * 2.f * xy + GS_C3_1 * SHK(10) * yz + GS_C3_2 * SHK(11) * -2.f * xy +
                                   GS_C3_3 * SHK(12) * -3.f * 2.f * xz + GS_C3_4 * SHK(13) * (-3.f * xx + 4.f * zz - yy) +
                                   GS_C3_5 * SHK(14) * 2.f * xz + GS_C3_6 * SHK(15) * 3.f * (xx - yy));
                            vy += (GS_C3_0 * SHK(9) * 3.f * (xx - yy) + GS_C3_1 * SHK(10) * xz + GS_C3_2 * SHK(11) * (-3.f * yy + 4.f * zz - xx) +
                                   GS_C3_3 * SHK(12) * -3.f * 2.f * yz + GS_C3_4 * SHK(13) * -2.f * xy +
                                   GS_C3_5 * SHK(14) * -2.f * yz + GS_C3_6 * SHK(15) * -3.f * 2.f * xy);
                            vz += (GS_C3_1 * SHK(10) * xy + GS_C3_2 * SHK(11) * 4.f * 2.f * yz + GS_C3_3 * SHK(12) * 3.f * (2.f * zz - xx - yy) +
                                   GS_C3_4 * SHK(13) * 4.f * 2.f * xz + GS_C3_5 * SHK(14) * (xx - yy));
                        }
                    }
#undef SHK
                    ax[ch] = vx; ay[ch] = vy; az[ch] = vz;
                }
                ddx = ax[0] * dR[0] + ax[1] * dR[1] + ax[2] * dR[2];
                ddy = ay[0] * dR[0] + ay[1] * dR[1] + ay[2] * dR[2];
                ddz = az[0] * dR[0] + az[1] * dR[1] + az[2] * dR[2];
            }
            // dnormvdv (auxiliary.h:107-117)
            const float sum2 = d0.x * d0.x + d0.y * d0.y + d0.z * d0.z;
            const float inv32 = 1.0f / sqrtf(sum2 * sum2 * sum2);
            dm3x += ((+sum2 - d0.x * d0.x) * ddx - d0.y * d0.x * ddy - d0.z * d0.x * ddz) * inv32;
            dm3y += (-d0.x * d0.y * ddx + (sum2 - d0.y * d0.y) * ddy - d0.z * d0.y * ddz) * inv32;
            dm3z += (-d0.x * d0.z * ddx - d0.y * d0.z * ddy + (sum2 - d0.z * d0.z) * ddz) * inv32;
        }

        // ---- cov3D backward (backward.cu:278-341)
        if (scales) {
            float R[3][3], M[3][3];
            gs_quat_R(q, R);
            const float sv[3] = {v.scale_modifier * sc.x, v.scale_modifier * sc.y, v.scale_modifier * sc.z};
#pragma unroll
            for (int r_ = 0; r_ < 3; r_++)
#pragma unroll
                for (int c_ = 0; c_ < 3; c_++) M[r_][c_] = sv[r_] * R[c_][r_];
            const float dS[3][3] = {{dcov[0], 0.5f * dcov[1], 0.5f * dcov[2]},
                                    {0.5f * dcov[1], dcov[3], 0.5f * dcov[4]},
                                    {0.5f * dcov[2], 0.5f * dcov[4], dcov[5]}};
            float dM[3][3];
#pragma unroll
            for (int r_ = 0; r_ < 3; r_++)
#pragma unroll
                for (int c_ = 0; c_ < 3; c_++)
                    dM[r_][c_] = 2.0f * (M[r_][0] * dS[0][c_] + M[r_][1] * dS[1][c_] + M[r_][2] * dS[2][c_]);
            dsx = R[0][0] * dM[0][0] + R[1][0] * dM[0][1] + R[2][0] * dM[0][2];
            dsy = R[0][1] * dM[1][0] + R[1][1] * dM[1][1] + R[2][1] * dM[1][2];
            dsz = R[0][2] * dM[2][0] + R[1][2] * dM[2][1] + R[2][2] * dM[2][2];
            float Q[3][3];
#pragma unroll
            for (int r_ = 0; r_ < 3; r_++)
#pragma unroll
                for (int c_ = 0; c_ < 3; c_++) Q[r_][c_] = dM[r_][c_] * sv[r_];
            const float r = q.x, x = q.y, y = q.z, z = q.w;
            drot.x = 2 * z * (Q[0][1] - Q[1][0]) + 2 * y * (Q[2][0] - Q[0][2]) + 2 * x * (Q[1][2] - Q[2][1]);
            drot.y = 2 * y * (Q[1][0] + Q[0][1]) + 2 * z * (Q[2][0] + Q[0][2]) + 2 * r * (Q[1][2] - Q[2][1]) - 4 * x * (Q[2][2] + Q[1][1]);
            drot.z = 2 * x * (Q[1][0] + Q[0][1]) + 2 * r * (Q[2][0] - Q[0][2]) + 2 * z * (Q[1][2] + Q[2][1]) - 4 * y * (Q[2][2] + Q[0][0]);
            drot.w = 2 * r * (Q[0][1] - Q[1][0]) + 2 * x * (Q[2][0] + Q[0][2]) + 2 * y * (Q[1][2] + Q[2][1]) - 4 * z * (Q[1][1] + Q[0][0]);
        }
    }

    // ------------------------------------------------------------------ coalesced stores
    float* st = s_stage[wid];
    if (g.dmeans3D) {
        st[3 * lane] = dm3x; st[3 * lane + 1] = dm3y; st[3 * lane + 2] = dm3z;
        __syncwarp();
        warp_store_rows<3>(g.dmeans3D, st, row0, P, lane);
        __syncwarp();
    }
    if (g.dmeans2D) {
        st[3 * lane] = dm2x; st[3 * lane + 1] = dm2y; st[3 * lane + 2] = 0.f;
        __syncwarp();
        warp_store_rows<3>(g.dmeans2D, st, row0, P, lane);
        __syncwarp();
    }
    if (g.dscales) {
        st[3 * lane] = dsx; st[3 * lane + 1] = dsy; st[3 * lane + 2] = dsz;
        __syncwarp();
        warp_store_rows<3>(g.dscales, st, row0, P, lane);
        __syncwarp();
    }
    if (g.dcolors) {
        st[3 * lane] = dcol0; st[3 * lane + 1] = dcol1; st[3 * lane + 2] = dcol2;
        __syncwarp();
        warp_store_rows<3>(g.dcolors, st, row0, P, lane);
        __syncwarp();
    }
    if (g.dcov3D) {
#pragma unroll
        for (int k = 0; k < 6; k++) st[6 * lane + k] = dcov[k];
        __syncwarp();
        warp_store_rows<6>(g.dcov3D, st, row0, P, lane);
        __syncwarp();
    }
    if (i < P) {
        if (g.dopacity) g.dopacity[i] = dop;
        if (g.drots) reinterpret_cast<float4*>(g.drots)[i] = drot;
    }
    if (g.dsh && v.M > 0) {
        // dL_dsh[g][k][ch] = basis_k(dir_g) * dL_dRGB_g[ch]: stage 19 floats per Gaussian, expand while storing
        float* sb = s_basis[wid];
        float* sr = s_rgb[wid];
#pragma unroll
        for (int k = 0; k < 16; k++) sb[lane * 16 + k] = bs[k];
        sr[lane * 3] = dRGB0; sr[lane * 3 + 1] = dRGB1; sr[lane * 3 + 2] = dRGB2;
        __syncwarp();
        const long long rows = min((long long)32, (long long)P - row0);
        if (rows > 0) {
            if (v.M == 16) store_dsh<48>(g.dsh + row0 * 48, sb, sr, (int)rows, lane);
            else store_dsh<0>(g.dsh + row0 * (v.M * 3), sb, sr, (int)rows, lane, v.M * 3);
        }
    }
}

}  // namespace

void gs_launch_gauss_bwd(const GsView& v, const int* radii, const float* means3D, const float* shs,
                         const float* scales, const float* rotations, const float* cov3D_precomp,
                         const float4* rec, float4* acc, GsGradPtrs g, cudaStream_t s) {
    const int grid = (v.P + kT - 1) / kT;
    k_gauss_bwd<<<grid, kT, 0, s>>>(v, radii, means3D, shs, scales, rotations, cov3D_precomp, rec, acc, g);
}

// dgr_backward.cuh — per-Gaussian backward (A6 of SURVEY.md §8a): takes the 10 moments the backward render
// reduced per Gaussian and applies the exact chain rule of the forward preprocess back to
// means3D / means2D / shs (or colors) / opacities / scales / rotations (or cov3D).
// Streaming, HBM-bound: re-reads 44 + 12 M bytes + 48 bytes of moments, writes 56 + 12 M bytes per Gaussian.
// Cull decisions, radius, tile rect, low-pass and eigenvalue floor are constants; the tx/tz, ty/tz clamp masks
// the gradient through the clamped ratio; SH channels clamped at 0 get no colour gradient.
#pragma once
#include "dgr_preprocess.cuh"

namespace dgr {

// d basis_k / d(x, y, z) for one coefficient index (k is a compile-time constant after unrolling)
__device__ __forceinline__ void sh_dbasis(int k, float x, float y, float z, float &dx, float &dy, float &dz) {
    const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
    dx = 0.f; dy = 0.f; dz = 0.f;
    switch (k) {
        case 1: dy = -DGR_SH_C1; break;
        case 2: dz = DGR_SH_C1; break;
        case 3: dx = -DGR_SH_C1; break;
        case 4: dx = DGR_SH_C2_0 * y; dy = DGR_SH_C2_0 * x; break;
        case 5: dy = DGR_SH_C2_1 * z; dz = DGR_SH_C2_1 * y; break;
        case 6: dx = DGR_SH_C2_2 * -2.f * x; dy = DGR_SH_C2_2 * -2.f * y; dz = DGR_SH_C2_2 * 4.f * z; break;
        case 7: dx = DGR_SH_C2_3 * z; dz = DGR_SH_C2_3 * x; break;
        case 8: dx = DGR_SH_C2_4 * 2.f * x; dy = DGR_SH_C2_4 * -2.f * y; break;
        case 9: dx = DGR_SH_C3_0 * 6.f * xy; dy = DGR_SH_C3_0 * (3.f * xx - 3.f * yy); break;
        case 10: dx = DGR_SH_C3_1 * yz; dy = DGR_SH_C3_1 * xz; dz = DGR_SH_C3_1 * xy; break;
        case 11: dx = DGR_SH_C3_2 * -2.f * xy; dy = DGR_SH_C3_2 * (4.f * zz - xx - 3.f * yy); dz = DGR_SH_C3_2 * 8.f * yz; break;
        case 12: dx = DGR_SH_C3_3 * -6.f * xz; dy = DGR_SH_C3_3 * -6.f * yz; dz = DGR_SH_C3_3 * (6.f * zz - 3.f * xx - 3.f * yy); break;
        case 13: dx = DGR_SH_C3_4 * (4.f * zz - 3.f * xx - yy); dy = DGR_SH_C3_4 * -2.f * xy; dz = DGR_SH_C3_4 * 8.f * xz; break;
        case 14: dx = DGR_SH_C3_5 * 2.f * xz; dy = DGR_SH_C3_5 * -2.f * yz; dz = DGR_SH_C3_5 * (xx - yy); break;
        case 15: dx = DGR_SH_C3_6 * (3.f * xx - 3.f * yy); dy = DGR_SH_C3_6 * -6.f * xy; break;
        default: break;
    }
}

// ---- output stage ---------------------------------------------------------------------------------------------------------------
// Every gradient segment is [P, K] row-major, so the 32 rows of a warp are ONE contiguous run of 32 K floats.  Each lane drops its
// K values into the warp's shared-memory image of that run and the warp writes the run back with 16-byte stores on consecutive
// addresses (one store instruction for a [32, 3] block instead of three strided ones; full 32-byte sectors, which is what matters
// when the destination is another GPU's memory behind NVLink).
//   acc:        add to what the LOCAL buffer holds (several views -> one gradient);
//   delta != 0: the result goes to the same element of this rank's slot in the owner rank's staging area (PeerPush) instead of
//               back to the local buffer.
constexpr int kOut3D = 0, kOut2D = 96, kOutOp = 192, kOutScale = 224, kOutRot = 320, kOutCov = 448, kOutCol = 640;    // float offsets per warp
constexpr int kShStrideVec = 28, kShStrideScalar = 25;
constexpr int kStageFloats = 32 * kShStrideVec;               // per warp: >= kOutCol + 32 * 3

__device__ __forceinline__ void warp_flush(float *seg, int K, long long delta, int row_base, int nrows, bool acc, const float *ws) {
    if (!seg) return;
    const int lane = threadIdx.x & 31;
    float *loc = seg + (size_t)row_base * K;
    float *dst = loc + delta;
    const int n = nrows * K;
    int done = 0;
    if ((((size_t)loc | (size_t)dst) & 15) == 0) {
        const int n4 = n >> 2;
        for (int i = lane; i < n4; i += 32) {
            float4 x = reinterpret_cast<const float4 *>(ws)[i];
            if (acc) { const float4 o = reinterpret_cast<const float4 *>(loc)[i]; x.x += o.x; x.y += o.y; x.z += o.z; x.w += o.w; }
            reinterpret_cast<float4 *>(dst)[i] = x;
        }
        done = n4 << 2;
    }
    for (int i = done + lane; i < n; i += 32) {
        float x = ws[i];
        if (acc) x += loc[i];
        dst[i] = x;
    }
}

// SH gradient rows of one warp's 32 Gaussians.  Each lane owns one row (3M floats, of which the first 3*NB are bs[k]*gr[ch] and the
// rest zero).  The rows go through shared memory in chunks of 24 floats; a chunk leaves as 16-byte stores, 96 contiguous bytes
// per row, when the row length is a multiple of 4 floats (row stride 28 words: conflict-free 16-byte accesses), else as 4-byte
// stores with consecutive lanes on consecutive addresses (row stride 25 words).
// FIRST = 1: the destination rows are GaussianModel._features_rest ([P, M-1, 3], coefficients 1 .. M-1).
template <int DEG, int FIRST = 0>
__device__ __forceinline__ void store_sh_grads(float *__restrict__ dL_dshs, int P, int M, const float (&bs)[16],
                                               const float (&gr)[3], bool acc, float *stage /* [kStageFloats] of this warp */,
                                               long long delta = 0) {
    constexpr int NB = (DEG + 1) * (DEG + 1) - FIRST;
    const int lane = threadIdx.x & 31;
    const int row_base = (int)(blockIdx.x * blockDim.x) + (int)(threadIdx.x & ~31u);
    const int nrows = min(32, P - row_base);
    if (nrows <= 0) return;
    const int rowlen = 3 * (M - FIRST);
    const bool vec = (rowlen & 3) == 0 && ((((size_t)dL_dshs) | ((size_t)(dL_dshs + delta))) & 15) == 0;
#pragma unroll
    for (int c = 0; c < 2; c++) {
        if (24 * c < rowlen) {
            const int cw = min(24, rowlen - 24 * c);
            if (vec) {
#pragma unroll
                for (int j4 = 0; j4 < 6; j4++) {
                    float v[4];
#pragma unroll
                    for (int e = 0; e < 4; e++) {
                        const int f = 24 * c + 4 * j4 + e;            // compile-time
                        v[e] = (f / 3 < NB) ? bs[(f / 3 + FIRST) < 16 ? (f / 3 + FIRST) : 0] * gr[f % 3] : 0.f;
                    }
                    if (4 * j4 < cw) reinterpret_cast<float4 *>(stage + lane * kShStrideVec)[j4] = make_float4(v[0], v[1], v[2], v[3]);
                }
                __syncwarp();
                const int q = cw >> 2, total = nrows * q;
                for (int i = lane; i < total; i += 32) {
                    const int row = i / q, c4 = i - row * q;
                    float4 x = *reinterpret_cast<const float4 *>(stage + row * kShStrideVec + 4 * c4);
                    float *loc = dL_dshs + (size_t)(row_base + row) * rowlen + 24 * c + 4 * c4;
                    if (acc) { const float4 o = *reinterpret_cast<const float4 *>(loc); x.x += o.x; x.y += o.y; x.z += o.z; x.w += o.w; }
                    *reinterpret_cast<float4 *>(loc + delta) = x;
                }
            } else {
#pragma unroll
                for (int j = 0; j < 24; j++) {
                    const int f = 24 * c + j;                 // compile-time
                    const float v = (f / 3 < NB) ? bs[(f / 3 + FIRST) < 16 ? (f / 3 + FIRST) : 0] * gr[f % 3] : 0.f;
                    if (f < rowlen) stage[lane * kShStrideScalar + j] = v;
                }
                __syncwarp();
                const unsigned inv = (65536u + (unsigned)cw - 1u) / (unsigned)cw;      // exact idx / cw for idx < 32 * 24
                for (int it = 0; it < cw; it++) {
                    const int idx = it * 32 + lane;
                    const int row = (int)(((unsigned)idx * inv) >> 16), col = idx - row * cw;
                    if (row < nrows) {
                        float *dst = dL_dshs + (size_t)(row_base + row) * rowlen + 24 * c + col;
                        const float v = stage[row * kShStrideScalar + col];
                        dst[delta] = acc ? (*dst + v) : v;
                    }
                }
            }
            __syncwarp();
        }
    }
}

// STAGE: the warp's inputs arrive by bulk TMA in its slice of dynamic shared memory (see dgr_preprocess.cuh, "Input staging");
// once every lane has its values in registers the same slice becomes the warp's output image (warp_flush), so the staged
// variant needs 8 x 32 x (44 + 12 M) bytes per CTA (59 KB at degree 3: three CTAs per SM, as before) and no static stage.
template <int DEG, bool HAS_SH, bool HAS_COV, bool RAW, bool STAGE = false>
__global__ void __launch_bounds__(kPreThreads, 3)
preprocess_bwd_kernel(int P, int M, int H, int W, float tanfovx, float tanfovy, float scale_modifier,
                      const float *__restrict__ viewmatrix, const float *__restrict__ projmatrix,
                      const float *__restrict__ campos,
                      const float *__restrict__ means3D, const float *__restrict__ shs, const float *__restrict__ shs_rest,
                      const float *__restrict__ opacities, const float *__restrict__ scales,
                      const float *__restrict__ rotations, const float *__restrict__ cov3D_precomp,
                      const int *__restrict__ radii, const unsigned *__restrict__ touched, const float *__restrict__ grad_rec,
                      float *__restrict__ dL_dmeans3D, float *__restrict__ dL_dmeans2D, float *__restrict__ dL_dshs,
                      float *__restrict__ dL_dcolors, float *__restrict__ dL_dopacities, float *__restrict__ dL_dscales,
                      float *__restrict__ dL_drotations, float *__restrict__ dL_dcov3D, float *__restrict__ dL_dshs_rest,
                      float *__restrict__ xyz_gradient_accum, float *__restrict__ denom, float *__restrict__ max_radii2D,
                      int accumulate, PeerPush push, int warp_smem_floats) {
    __shared__ FrameConsts fc;
    __shared__ __align__(16) float s_stage[STAGE ? 4 : (kPreThreads / 32) * kStageFloats];
    __shared__ uint64_t s_stage_bar[kPreThreads / 32];
    extern __shared__ __align__(128) float s_dyn[];              // STAGE: 8 warps x warp_smem_floats (>= kStageFloats)
    static_assert(!STAGE || (HAS_SH && !HAS_COV), "input staging: SH + scale / rotation inputs only");
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    bool staged = false;                                         // warp-uniform: all 32 Gaussians of this warp exist and are staged
    if (STAGE) {
        const int g0 = (int)(blockIdx.x * blockDim.x + (threadIdx.x & ~31u));
        uint64_t *wbar = &s_stage_bar[threadIdx.x >> 5];
        staged = g0 + 32 <= P;
        if ((threadIdx.x & 31) == 0) {
            mbar_init(wbar, 1); mbar_fence_init();
            if (staged) stage_issue<RAW>(s_dyn + (threadIdx.x >> 5) * warp_smem_floats, wbar, g0, M, means3D, scales, rotations, opacities, shs, shs_rest);
        }
        __syncwarp();
    }
    load_frame(fc, viewmatrix, projmatrix, HAS_SH ? campos : nullptr);
    __syncthreads();
    const bool acc = accumulate != 0;
    const bool in_range = g < P;
    // gradient rows of Gaussians another rank owns go to that rank's staging slot (what the local buffer has accumulated over
    // this rank's earlier views of the iteration + this view); every row is written then, culled Gaussians included
    const long long delta = push.per > 0 ? push.delta[(blockIdx.x * blockDim.x) / (unsigned)push.per] : 0;
    const bool pushing = delta != 0;
    float *ws = STAGE ? s_dyn + (threadIdx.x >> 5) * warp_smem_floats
                      : s_stage + (threadIdx.x >> 5) * kStageFloats;   // this warp's output image (see warp_flush)
    const int lane = threadIdx.x & 31;
    // every input of this Gaussian is requested up front (one memory round trip; none of it depends on the backward render)
    int radius_in = 0;
    unsigned touched_in = 0;
    float3 p = make_float3(0.f, 0.f, 0.f), s_in = make_float3(0.f, 0.f, 0.f);
    float4 q_in = make_float4(1.f, 0.f, 0.f, 0.f);
    float o_in = 0.f, S6[6], shc[48];
    if (in_range) {
        radius_in = radii[g];
        touched_in = __ldg(touched + g);
    }
    if (staged) {
        mbar_wait(&s_stage_bar[threadIdx.x >> 5], 0u);
        p = make_float3(ws[kStgMean + 3 * lane], ws[kStgMean + 3 * lane + 1], ws[kStgMean + 3 * lane + 2]);
        s_in = make_float3(ws[kStgScale + 3 * lane], ws[kStgScale + 3 * lane + 1], ws[kStgScale + 3 * lane + 2]);
        q_in = *reinterpret_cast<const float4 *>(ws + kStgRot + 4 * lane);
        o_in = ws[kStgOp + lane];
        if (RAW) { if (DEG > 0) lds_sh_row<DEG, 1>(ws + kStgSh + 96 + lane * (M - 1) * 3, ((M - 1) & 3) == 0, shc); }
        else lds_sh_row<DEG>(ws + kStgSh + lane * M * 3, (M & 3) == 0, shc);
    }
    if (STAGE) __syncwarp();             // every lane has its inputs in registers: from here on the slice is the warp's output image
    if (in_range && !staged) {
        p = make_float3(__ldg(means3D + 3 * (size_t)g), __ldg(means3D + 3 * (size_t)g + 1), __ldg(means3D + 3 * (size_t)g + 2));
        if (HAS_COV) {
#pragma unroll
            for (int i = 0; i < 6; i++) S6[i] = __ldg(cov3D_precomp + 6 * (size_t)g + i);
        } else {
            s_in = make_float3(__ldg(scales + 3 * (size_t)g), __ldg(scales + 3 * (size_t)g + 1), __ldg(scales + 3 * (size_t)g + 2));
            q_in = ldg_f4(rotations + 4 * (size_t)g);
        }
        o_in = __ldg(opacities + g);
        if (HAS_SH) {
            if (RAW) { if (DEG > 0) load_sh_row<DEG, 1>(shs_rest + (size_t)g * (M - 1) * 3, ((M - 1) & 3) == 0, shc); }       // d basis_0 = 0
            else load_sh_row<DEG>(shs + (size_t)g * M * 3, (M & 3) == 0, shc);
        }
    }
    const bool visible = in_range && radius_in > 0;
    float bs[16];
    float gr[3] = { 0.f, 0.f, 0.f };
#pragma unroll
    for (int k = 0; k < 16; k++) bs[k] = 0.f;

    if (!visible) {                                                 // culled (or past the end): a zero row in every segment
#pragma unroll
        for (int k = 0; k < 3; k++) { ws[kOut3D + 3 * lane + k] = 0.f; ws[kOut2D + 3 * lane + k] = 0.f; ws[kOutScale + 3 * lane + k] = 0.f; }
        ws[kOutOp + lane] = 0.f;
        *reinterpret_cast<float4 *>(ws + kOutRot + 4 * lane) = make_float4(0.f, 0.f, 0.f, 0.f);
        if (HAS_COV) { for (int k = 0; k < 6; k++) ws[kOutCov + 6 * lane + k] = 0.f; }
        if (!HAS_SH) { for (int k = 0; k < 3; k++) ws[kOutCol + 3 * lane + k] = 0.f; }
    }
    if (visible) {
        float R[9];
        float3 s = make_float3(0.f, 0.f, 0.f);
        float4 q = make_float4(1.f, 0.f, 0.f, 0.f);
        float inv_qnorm = 1.f;
        if (!HAS_COV) {
            s = s_in;
            if (RAW) s = make_float3(expf(s.x), expf(s.y), expf(s.z));
            s = make_float3(scale_modifier * s.x, scale_modifier * s.y, scale_modifier * s.z);
            q = q_in;
            if (RAW) q = act_normalize(q, inv_qnorm);
            quat_to_R(q, R);
            cov3d_from_scale_rot(s, R, S6);
        }
        const float fx = (float)W / (2.f * tanfovx), fy = (float)H / (2.f * tanfovy);
        Geo geo;
        project_geo(fc, p, S6, fx, fy, DGR_FOV_CLAMP * tanfovx, DGR_FOV_CLAMP * tanfovy, geo);
        const float a = geo.cxx, b = geo.cxy, c = geo.cyy;
        const float det = a * c - b * b, di = 1.f / det;
        const float cA = c * di, cB = -b * di, cC = a * di;
        const float o = RAW ? act_sigmoid(o_in) : o_in;

        // Everything above reads only the op's inputs: under a programmatic dependent launch it overlaps the tail of the
        // backward render kernel.  The moments that kernel accumulates are complete after this point.
        pdl_wait();
        const float4 m0 = ldg_f4(grad_rec + (size_t)g * kGradRecFloats);
        const float4 m1 = ldg_f4(grad_rec + (size_t)g * kGradRecFloats + 4);
        const float4 m2 = ldg_f4(grad_rec + (size_t)g * kGradRecFloats + 8);
        // moments: m0 = {u, u dx, u dy, u dx^2}, m1 = {u dx dy, u dy^2, wr, wg}, m2 = {wb, wd, -, -}

        // pixel-space mean gradient and conic gradient from the moments
        const float gpx = -(cA * m0.y + cB * m0.z), gpy = -(cC * m0.z + cB * m0.y);
        const float gA = -0.5f * m0.w, gB = -m1.x, gC = -0.5f * m1.y;
        const float g_op = (o != 0.f) ? m0.x / o : 0.f;
        const float g_rgb[3] = { m1.z, m1.w, m2.x };
        const float g_depth = m2.y;

        float dmean[3] = { 0.f, 0.f, 0.f };
        // (1) mean_px through the full projection; reported means2D grad is in NDC units
        const float gndx = gpx * 0.5f * (float)W, gndy = gpy * 0.5f * (float)H;
        ws[kOut2D + 3 * lane] = gndx; ws[kOut2D + 3 * lane + 1] = gndy; ws[kOut2D + 3 * lane + 2] = 0.f;
        // densification bookkeeping of the training loop (gs_renderer.py:625-627, main.py:279-281), for visible Gaussians:
        // xyz_gradient_accum += |d L / d means2D[:, :2]| of THIS render, denom += 1, max_radii2D = max(max_radii2D, radii)
        if (xyz_gradient_accum) xyz_gradient_accum[g] += sqrtf(gndx * gndx + gndy * gndy);
        if (denom) denom[g] += 1.f;
        if (max_radii2D) max_radii2D[g] = fmaxf(max_radii2D[g], (float)radius_in);
        {
            const float *PM = fc.PM;
            const float mul1 = geo.ndcx * geo.pw, mul2 = geo.ndcy * geo.pw;     // ph.x * pw^2, ph.y * pw^2
#pragma unroll
            for (int k = 0; k < 3; k++)
                dmean[k] += (PM[4 * k] * geo.pw - PM[4 * k + 3] * mul1) * gndx + (PM[4 * k + 1] * geo.pw - PM[4 * k + 3] * mul2) * gndy;
        }
        // (4) depth = t.z
#pragma unroll
        for (int k = 0; k < 3; k++) dmean[k] += fc.V[4 * k + 2] * g_depth;
        // (3) colour
        if (HAS_SH) {
            float dx = p.x - fc.cam[0], dy = p.y - fc.cam[1], dz = p.z - fc.cam[2];
            const float il = 1.f / sqrtf(dx * dx + dy * dy + dz * dz);
            dx *= il; dy *= il; dz *= il;
            sh_basis<DEG>(dx, dy, dz, bs);
            const unsigned flags = touched_in >> 29;                     // SH channels the forward clamped at 0
#pragma unroll
            for (int ch = 0; ch < 3; ch++) gr[ch] = ((flags >> ch) & 1u) ? 0.f : g_rgb[ch];
            float ddx = 0.f, ddy = 0.f, ddz = 0.f;
            constexpr int NB = (DEG + 1) * (DEG + 1);
#pragma unroll
            for (int k = 1; k < NB; k++) {                                // d basis_0 = 0
                const int off = RAW ? 3 * (k - 1) : 3 * k;
                const float dotc = shc[off] * gr[0] + shc[off + 1] * gr[1] + shc[off + 2] * gr[2];
                float bx, by, bz;
                sh_dbasis(k, dx, dy, dz, bx, by, bz);
                ddx += bx * dotc; ddy += by * dotc; ddz += bz * dotc;
            }
            const float dot = dx * ddx + dy * ddy + dz * ddz;
            dmean[0] += (ddx - dx * dot) * il; dmean[1] += (ddy - dy * dot) * il; dmean[2] += (ddz - dz * dot) * il;
        } else {
#pragma unroll
            for (int ch = 0; ch < 3; ch++) ws[kOutCol + 3 * lane + ch] = g_rgb[ch];
        }
        ws[kOutOp + lane] = RAW ? g_op * o * (1.f - o) : g_op;                                              // sigmoid'

        // (2) conic -> cov2D -> (Sigma, T = J Rwv)
        const float d2 = di * di;
        const float ga = d2 * (-c * c * gA + b * c * gB - b * b * gC);
        const float gc = d2 * (-b * b * gA + a * b * gB - a * a * gC);
        const float gb = d2 * (2.f * b * c * gA - (det + 2.f * b * b) * gB + 2.f * a * b * gC);
        const float G00 = ga, G01 = 0.5f * gb, G11 = gc;
        const float *T0 = geo.T0, *T1 = geo.T1;
        float dS[9];
#pragma unroll
        for (int k = 0; k < 3; k++)
#pragma unroll
            for (int l = 0; l < 3; l++)
                dS[3 * k + l] = T0[k] * (G00 * T0[l] + G01 * T1[l]) + T1[k] * (G01 * T0[l] + G11 * T1[l]);
        const float Sg[9] = { S6[0], S6[1], S6[2], S6[1], S6[3], S6[4], S6[2], S6[4], S6[5] };
        float TS0[3], TS1[3];
#pragma unroll
        for (int k = 0; k < 3; k++) {
            TS0[k] = T0[0] * Sg[k] + T0[1] * Sg[3 + k] + T0[2] * Sg[6 + k];
            TS1[k] = T1[0] * Sg[k] + T1[1] * Sg[3 + k] + T1[2] * Sg[6 + k];
        }
        float dT0[3], dT1[3];
#pragma unroll
        for (int k = 0; k < 3; k++) { dT0[k] = 2.f * (G00 * TS0[k] + G01 * TS1[k]); dT1[k] = 2.f * (G01 * TS0[k] + G11 * TS1[k]); }
        float dJ00 = 0.f, dJ02 = 0.f, dJ11 = 0.f, dJ12 = 0.f;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            dJ00 += dT0[k] * fc.V[4 * k]; dJ02 += dT0[k] * fc.V[4 * k + 2];
            dJ11 += dT1[k] * fc.V[4 * k + 1]; dJ12 += dT1[k] * fc.V[4 * k + 2];
        }
        const float tz = geo.t[2];
        const float tz2 = 1.f / (tz * tz), tz3 = tz2 / tz;
        const float dtx = geo.clx ? 0.f : -fx * tz2 * dJ02;
        const float dty = geo.cly ? 0.f : -fy * tz2 * dJ12;
        const float dtz = -fx * tz2 * dJ00 - fy * tz2 * dJ11 + 2.f * fx * geo.tx * tz3 * dJ02 + 2.f * fy * geo.ty * tz3 * dJ12;
#pragma unroll
        for (int k = 0; k < 3; k++) dmean[k] += fc.V[4 * k] * dtx + fc.V[4 * k + 1] * dty + fc.V[4 * k + 2] * dtz;
#pragma unroll
        for (int k = 0; k < 3; k++) ws[kOut3D + 3 * lane + k] = dmean[k];
        if (HAS_COV) {
            float *out = ws + kOutCov + 6 * lane;
            out[0] = dS[0]; out[1] = 2.f * dS[1]; out[2] = 2.f * dS[2]; out[3] = dS[4]; out[4] = 2.f * dS[5]; out[5] = dS[8];
        } else {
            // Sigma = Mx Mx^T, Mx = R diag(s):  dL/dMx = 2 dS Mx
            float Mx[9];
#pragma unroll
            for (int i = 0; i < 3; i++) { Mx[3 * i] = R[3 * i] * s.x; Mx[3 * i + 1] = R[3 * i + 1] * s.y; Mx[3 * i + 2] = R[3 * i + 2] * s.z; }
            float dM[9];
#pragma unroll
            for (int i = 0; i < 3; i++)
#pragma unroll
                for (int k = 0; k < 3; k++)
                    dM[3 * i + k] = 2.f * (dS[3 * i] * Mx[k] + dS[3 * i + 1] * Mx[3 + k] + dS[3 * i + 2] * Mx[6 + k]);
            const float sv[3] = { s.x, s.y, s.z };
            float dR[9];
#pragma unroll
            for (int k = 0; k < 3; k++) {
                float v = 0.f;
#pragma unroll
                for (int i = 0; i < 3; i++) { v += dM[3 * i + k] * R[3 * i + k]; dR[3 * i + k] = dM[3 * i + k] * sv[k]; }
                ws[kOutScale + 3 * lane + k] = RAW ? v * sv[k] : v * scale_modifier;                                   // exp' = exp
            }
            {
                const float r = q.x, x = q.y, y = q.z, z = q.w;
                float dq[4];
                dq[0] = 2.f * (-z * dR[1] + y * dR[2] + z * dR[3] - x * dR[5] - y * dR[6] + x * dR[7]);
                dq[1] = 2.f * (y * dR[1] + z * dR[2] + y * dR[3] - 2.f * x * dR[4] - r * dR[5] + z * dR[6] + r * dR[7] - 2.f * x * dR[8]);
                dq[2] = 2.f * (-2.f * y * dR[0] + x * dR[1] + r * dR[2] + x * dR[3] + z * dR[5] - r * dR[6] + z * dR[7] - 2.f * y * dR[8]);
                dq[3] = 2.f * (-2.f * z * dR[0] - r * dR[1] + x * dR[2] + r * dR[3] - 2.f * z * dR[4] + y * dR[5] + x * dR[6] + y * dR[7]);
                if (RAW) {          // q = raw / |raw|:  d raw = (dq - q (q . dq)) / |raw|
                    const float qd = r * dq[0] + x * dq[1] + y * dq[2] + z * dq[3];
                    dq[0] = (dq[0] - r * qd) * inv_qnorm; dq[1] = (dq[1] - x * qd) * inv_qnorm;
                    dq[2] = (dq[2] - y * qd) * inv_qnorm; dq[3] = (dq[3] - z * qd) * inv_qnorm;
                }
                *reinterpret_cast<float4 *>(ws + kOutRot + 4 * lane) = make_float4(dq[0], dq[1], dq[2], dq[3]);
            }
        }
    }
    // ---- the warp's rows leave together (see warp_flush); rows past the end of the arrays are not written
    __syncwarp();
    {
        const int row_base = (int)(blockIdx.x * blockDim.x) + (int)(threadIdx.x & ~31u);
        const int nrows = min(32, P - row_base);
        if (nrows > 0) {
            warp_flush(dL_dmeans3D, 3, delta, row_base, nrows, acc, ws + kOut3D);
            warp_flush(dL_dmeans2D, 3, delta, row_base, nrows, acc, ws + kOut2D);
            warp_flush(dL_dopacities, 1, delta, row_base, nrows, acc, ws + kOutOp);
            if (!HAS_COV) {
                warp_flush(dL_dscales, 3, delta, row_base, nrows, acc, ws + kOutScale);
                warp_flush(dL_drotations, 4, delta, row_base, nrows, acc, ws + kOutRot);
            } else warp_flush(dL_dcov3D, 6, delta, row_base, nrows, acc, ws + kOutCov);
            if (!HAS_SH) warp_flush(dL_dcolors, 3, delta, row_base, nrows, acc, ws + kOutCol);
        }
        __syncwarp();
        if (HAS_SH && RAW) {
            if (dL_dshs && nrows > 0) {           // _features_dc gradient [P,1,3] (zero for culled Gaussians: bs = gr = 0)
#pragma unroll
                for (int ch = 0; ch < 3; ch++) ws[3 * lane + ch] = bs[0] * gr[ch];
                __syncwarp();
                warp_flush(dL_dshs, 3, delta, row_base, nrows, acc, ws);
                __syncwarp();
            }
            if (dL_dshs_rest && M > 1) store_sh_grads<DEG, 1>(dL_dshs_rest, P, M, bs, gr, acc, ws, delta);
        } else if (HAS_SH && dL_dshs) store_sh_grads<DEG>(dL_dshs, P, M, bs, gr, acc, ws, delta);
    }
    if (pushing) __threadfence_system();          // the rows are in the owner's memory when this grid completes
}

}  // namespace dgr

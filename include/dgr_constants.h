/*
 * dgr_constants.h — every threshold of the Gaussian-splat rasterizer path in ONE place.
 *
 * All of these are recalled from the public diff-gaussian-rasterization behaviour (ashawkey fork,
 * depth+alpha outputs); that package is NOT vendored in the reference, so each value is
 * tagged UNVERIFIED-EXT: a later correction is a one-line change here, picked up by the CUDA kernels
 * (dreamgaussian_b200/csrc) and by the CPU oracle (oracle/dgr_oracle.c) alike.
 *
 * In-tree anchors that DO pin values: SH constants = /root/reference/sh_utils.py:26-43,
 * SH offset +0.5 / clamp >= 0 = /root/reference/gs_renderer.py:793.
 */
#ifndef DGR_CONSTANTS_H
#define DGR_CONSTANTS_H

#define DGR_TILE            16          /* UNVERIFIED-EXT tile edge in pixels (BLOCK_X = BLOCK_Y)            */
#define DGR_NEAR_CULL_Z     0.2f        /* UNVERIFIED-EXT view-space z at or below which a Gaussian is culled */
#define DGR_W_EPS           1e-7f       /* UNVERIFIED-EXT added to clip-space w before the divide             */
#define DGR_COV2D_LOWPASS   0.3f        /* UNVERIFIED-EXT added to cov2D xx, yy                               */
#define DGR_EIG_FLOOR       0.1f        /* UNVERIFIED-EXT floor under the eigenvalue discriminant             */
#define DGR_RADIUS_SIGMAS   3.0f        /* UNVERIFIED-EXT radius = ceil(3 sqrt(lambda_max))                   */
#define DGR_FOV_CLAMP       1.3f        /* UNVERIFIED-EXT tx/tz, ty/tz clamped to +-1.3 tan(fov/2) inside J   */
#define DGR_ALPHA_MAX       0.99f       /* UNVERIFIED-EXT alpha = min(0.99, o * G)                            */
#define DGR_ALPHA_MIN       (1.0f / 255.0f) /* UNVERIFIED-EXT contributions below this are skipped           */
#define DGR_T_STOP          1e-4f       /* UNVERIFIED-EXT stop when T * (1 - alpha) < 1e-4                    */
#define DGR_SH_OFFSET       0.5f        /* gs_renderer.py:793                                                 */

/* sh_utils.py:26-43 */
#define DGR_SH_C0   0.28209479177387814f
#define DGR_SH_C1   0.4886025119029199f
#define DGR_SH_C2_0 1.0925484305920792f
#define DGR_SH_C2_1 -1.0925484305920792f
#define DGR_SH_C2_2 0.31539156525252005f
#define DGR_SH_C2_3 -1.0925484305920792f
#define DGR_SH_C2_4 0.5462742152960396f
#define DGR_SH_C3_0 -0.5900435899266435f
#define DGR_SH_C3_1 2.890611442640554f
#define DGR_SH_C3_2 -0.4570457994644658f
#define DGR_SH_C3_3 0.3731763325901154f
#define DGR_SH_C3_4 -0.4570457994644658f
#define DGR_SH_C3_5 1.445305721320277f
#define DGR_SH_C3_6 -0.5900435899266435f

#endif /* DGR_CONSTANTS_H */

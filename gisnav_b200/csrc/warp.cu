// warp.cu — StereoNode's rotate + centre-crop of the orthoimage/DEM stack as one device kernel
// (SURVEY.md §8(f) rank 2): the step immediately in front of the extractor, so the rotated raster
// and DEM can stay in HBM and feed K1 / K5 without a host round-trip.
//
// Replaces, fused into a single pass over the crop window:
//   cv2.cvtColor(orthoimage, COLOR_BGR2GRAY)                 ros/gisnav/gisnav/core/stereo_node.py:239
//   cv2.getRotationMatrix2D + cv2.warpAffine(stack, M, (w,h)) stereo_node.py:313-316
//   rotated[dy:dy+H, dx:dx+W]                                 stereo_node.py:319-323
// Only the cropped window is evaluated (the reference warps the whole padded square and throws the
// rim away).  The arithmetic is OpenCV's fixed-point path restated (oracle/stereo_ref.py, pinned
// bit-for-bit against the installed cv2): float64 for the per-row / per-column offsets (explicit
// _rn intrinsics, no FMA contraction), then integers only.  Result: bit-identical to the reference.
//
// Roofline: HBM-bound byte work.  Algorithmic bytes per output pixel: (channels + 1) read (the
// source footprint of a rotation equals the crop area) + 2 written.
#include "common.cuh"

#include <math.h>

#define WP_TILE_W 128
#define WP_TILE_H 16
#define WP_THREADS 256

struct WarpMat { double m[6]; };  // inverse map: destination (x,y) -> source, as cv::warpAffine uses it

__device__ __forceinline__ int wp_gray(const uint8_t* __restrict__ p) {
    // BGR -> Y, 15-bit coefficients (cv::cvtColor u8)
    return (int)((__ldg(p) * 3735u + __ldg(p + 1) * 19235u + __ldg(p + 2) * 9798u + (1u << 14)) >> 15);
}

template <int CH>
__device__ __forceinline__ int wp_fetch(const uint8_t* __restrict__ src, int h, int w, int y, int x) {
    if ((unsigned)x >= (unsigned)w || (unsigned)y >= (unsigned)h) return 0;  // BORDER_CONSTANT, value 0
    const uint8_t* p = src + ((size_t)y * w + x) * CH;
    if (CH == 3) return wp_gray(p);
    return (int)__ldg(p);
}

template <int CH>
__global__ void __launch_bounds__(WP_THREADS) rotate_crop_kernel(const uint8_t* __restrict__ ortho, const uint8_t* __restrict__ dem,
                                                                  int h, int w, WarpMat mat, int dx, int dy, int crop_h, int crop_w,
                                                                  uint8_t* __restrict__ out_ref, uint8_t* __restrict__ out_dem) {
    __shared__ int s_ad[WP_TILE_W], s_bd[WP_TILE_W], s_x0[WP_TILE_H], s_y0[WP_TILE_H];
    const int tx0 = blockIdx.x * WP_TILE_W, ty0 = blockIdx.y * WP_TILE_H;
    const int t = threadIdx.x;
    if (t < WP_TILE_W) {
        // adelta[x] = saturate_cast<int>(M[0]*x*AB_SCALE), bdelta[x] = saturate_cast<int>(M[3]*x*AB_SCALE)
        const double x = (double)(tx0 + t + dx);
        s_ad[t] = __double2int_rn(__dmul_rn(__dmul_rn(mat.m[0], x), 1024.0));
        s_bd[t] = __double2int_rn(__dmul_rn(__dmul_rn(mat.m[3], x), 1024.0));
    } else if (t < WP_TILE_W + WP_TILE_H) {
        // X0 = saturate_cast<int>((M[1]*y + M[2])*AB_SCALE) + round_delta (= AB_SCALE / INTER_TAB_SIZE / 2 = 16)
        const int r = t - WP_TILE_W;
        const double y = (double)(ty0 + r + dy);
        s_x0[r] = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(mat.m[1], y), mat.m[2]), 1024.0)) + 16;
        s_y0[r] = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(mat.m[4], y), mat.m[5]), 1024.0)) + 16;
    }
    __syncthreads();
    const int cg = t & 31;          // group of 4 consecutive columns
    const bool vec = ((crop_w & 3) == 0) && ((((uintptr_t)out_ref | (uintptr_t)out_dem) & 3) == 0);
    for (int r = t >> 5; r < WP_TILE_H; r += WP_THREADS / 32) {
        const int oy = ty0 + r;
        if (oy >= crop_h) break;
        uint32_t pack_ref = 0, pack_dem = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int lx = cg * 4 + k, ox = tx0 + lx;
            int vr = 0, vd = 0;
            if (ox < crop_w) {
                const int X = (s_x0[r] + s_ad[lx]) >> 5, Y = (s_y0[r] + s_bd[lx]) >> 5;   // AB_BITS - INTER_BITS
                int sx = X >> 5, sy = Y >> 5;
                sx = min(max(sx, -32768), 32767);   // saturate_cast<short>
                sy = min(max(sy, -32768), 32767);
                const int fx = X & 31, fy = Y & 31;
                const int w00 = (32 - fx) * (32 - fy) * 32, w01 = fx * (32 - fy) * 32, w10 = (32 - fx) * fy * 32, w11 = fx * fy * 32;
                const int a = wp_fetch<CH>(ortho, h, w, sy, sx) * w00 + wp_fetch<CH>(ortho, h, w, sy, sx + 1) * w01 +
                              wp_fetch<CH>(ortho, h, w, sy + 1, sx) * w10 + wp_fetch<CH>(ortho, h, w, sy + 1, sx + 1) * w11;
                vr = min((a + (1 << 14)) >> 15, 255);
                if (dem) {
                    const int d = wp_fetch<1>(dem, h, w, sy, sx) * w00 + wp_fetch<1>(dem, h, w, sy, sx + 1) * w01 +
                                  wp_fetch<1>(dem, h, w, sy + 1, sx) * w10 + wp_fetch<1>(dem, h, w, sy + 1, sx + 1) * w11;
                    vd = min((d + (1 << 14)) >> 15, 255);
                }
                if (!vec) {
                    out_ref[(size_t)oy * crop_w + ox] = (uint8_t)vr;
                    if (dem) out_dem[(size_t)oy * crop_w + ox] = (uint8_t)vd;
                }
            }
            pack_ref |= (uint32_t)vr << (8 * k);
            pack_dem |= (uint32_t)vd << (8 * k);
        }
        const int ox0 = tx0 + cg * 4;
        if (vec && ox0 < crop_w) {
            *reinterpret_cast<uint32_t*>(out_ref + (size_t)oy * crop_w + ox0) = pack_ref;
            if (dem) *reinterpret_cast<uint32_t*>(out_dem + (size_t)oy * crop_w + ox0) = pack_dem;
        }
    }
}

// cv::getRotationMatrix2D(center, angle, 1.0) followed by the forward->inverse step of cv::warpAffine
static void rotation_and_inverse(int h, int w, double angle_degrees, double fwd[6], double inv[6]) {
    const double cx = (double)(w / 2), cy = (double)(h / 2);
    const double angle = angle_degrees * (3.1415926535897932384626433832795 / 180.0);  // angle *= CV_PI/180
    const double a = cos(angle) * 1.0, b = sin(angle) * 1.0;
    fwd[0] = a; fwd[1] = b; fwd[2] = (1 - a) * cx - b * cy;
    fwd[3] = -b; fwd[4] = a; fwd[5] = b * cx + (1 - a) * cy;
    double m[6];
    for (int i = 0; i < 6; ++i) m[i] = fwd[i];
    double d = m[0] * m[4] - m[1] * m[3];
    d = d != 0 ? 1. / d : 0;
    const double a11 = m[4] * d, a22 = m[0] * d;
    m[0] = a11; m[1] *= -d;
    m[3] *= -d; m[4] = a22;
    const double b1 = -m[0] * m[2] - m[1] * m[5];
    const double b2 = -m[3] * m[2] - m[4] * m[5];
    m[2] = b1; m[5] = b2;
    for (int i = 0; i < 6; ++i) inv[i] = m[i];
}

static int ensure_warp_buf(gnb_ctx* ctx, size_t bytes) {
    if (bytes <= ctx->warp_bytes) return GNB_OK;
    if (ctx->warp_buf) cudaFree(ctx->warp_buf);
    ctx->warp_buf = nullptr; ctx->warp_bytes = 0;
    GNB_CUDA(ctx, cudaMalloc((void**)&ctx->warp_buf, bytes));
    ctx->warp_bytes = bytes;
    return GNB_OK;
}

extern "C" int gnb_rotate_crop(gnb_ctx* ctx, const uint8_t* ortho, int channels, const uint8_t* dem, int h, int w,
                               double angle_degrees, int crop_h, int crop_w, int on_device, uint8_t* out_ref, uint8_t* out_dem,
                               double* out_rotation6, double* out_inverse9) {
    if (!ctx || !ortho || !out_ref || (dem && !out_dem)) return GNB_E_INVALID;
    GNB_CUDA(ctx, cudaSetDevice(ctx->device));
    if ((channels != 1 && channels != 3) || h < 1 || w < 1 || crop_h < 1 || crop_w < 1) {
        GNB_SET_ERR(ctx, "gnb_rotate_crop: channels must be 1 (gray) or 3 (BGR), sizes positive");
        return GNB_E_INVALID;
    }
    // crop window inside the rotated canvas (stereo_node.py:319-323); the reference's slice silently
    // misbehaves for a negative origin, here it is a usage error
    const int dx = w / 2 - crop_w / 2, dy = h / 2 - crop_h / 2;
    if (dx < 0 || dy < 0 || dx + crop_w > w || dy + crop_h > h) {
        GNB_SET_ERR(ctx, "gnb_rotate_crop: crop %dx%d does not fit the %dx%d orthoimage", crop_h, crop_w, h, w);
        return GNB_E_INVALID;
    }
    double fwd[6], inv[6];
    rotation_and_inverse(h, w, angle_degrees, fwd, inv);
    if (out_rotation6) memcpy(out_rotation6, fwd, sizeof(fwd));
    if (out_inverse9) {
        // inverse of [fwd; 0 0 1] times the crop translation (stereo_node.py:326-333), closed form
        const double i9[9] = {inv[0], inv[1], inv[0] * dx + inv[1] * dy + inv[2],
                              inv[3], inv[4], inv[3] * dx + inv[4] * dy + inv[5], 0, 0, 1};
        memcpy(out_inverse9, i9, sizeof(i9));
    }
    const size_t src_px = (size_t)h * w, crop_px = (size_t)crop_h * crop_w;
    const uint8_t *d_ortho = ortho, *d_dem = dem;
    uint8_t *d_ref = out_ref, *d_odem = out_dem;
    if (!on_device) {
        // layout of the staging buffer: [ortho | dem | out_ref | out_dem], each 16-byte aligned
        auto al = [](size_t x) { return (x + 15) & ~(size_t)15; };
        const size_t o_dem = al(src_px * channels), o_ref = o_dem + al(dem ? src_px : 0), o_odem = o_ref + al(crop_px);
        int rc;
        if ((rc = ensure_warp_buf(ctx, o_odem + al(crop_px)))) return rc;
        GNB_CUDA(ctx, cudaMemcpyAsync(ctx->warp_buf, ortho, src_px * channels, cudaMemcpyHostToDevice, ctx->stream));
        if (dem) GNB_CUDA(ctx, cudaMemcpyAsync(ctx->warp_buf + o_dem, dem, src_px, cudaMemcpyHostToDevice, ctx->stream));
        d_ortho = ctx->warp_buf; d_dem = dem ? ctx->warp_buf + o_dem : nullptr;
        d_ref = ctx->warp_buf + o_ref; d_odem = ctx->warp_buf + o_odem;
    }
    WarpMat mat;
    for (int i = 0; i < 6; ++i) mat.m[i] = inv[i];
    const dim3 grid(ceil_div(crop_w, WP_TILE_W), ceil_div(crop_h, WP_TILE_H));
    if (channels == 3) {
        GNB_KERNEL(ctx, "rotate_crop_kernel", rotate_crop_kernel<3><<<grid, WP_THREADS, 0, ctx->stream>>>(
            d_ortho, d_dem, h, w, mat, dx, dy, crop_h, crop_w, d_ref, d_odem));
    } else {
        GNB_KERNEL(ctx, "rotate_crop_kernel", rotate_crop_kernel<1><<<grid, WP_THREADS, 0, ctx->stream>>>(
            d_ortho, d_dem, h, w, mat, dx, dy, crop_h, crop_w, d_ref, d_odem));
    }
    if (!on_device) {
        GNB_CUDA(ctx, cudaMemcpyAsync(out_ref, d_ref, crop_px, cudaMemcpyDeviceToHost, ctx->stream));
        if (dem) GNB_CUDA(ctx, cudaMemcpyAsync(out_dem, d_odem, crop_px, cudaMemcpyDeviceToHost, ctx->stream));
    }
    GNB_SYNC(ctx);
    return GNB_OK;
}

// bayesn_tiles.cuh - per-SN band-weight tiles built on the device
// (reference SEDmodel._calculate_band_weights, bayesn/bayesn_model.py:499-554, plus the packing of
// bayesn_b200/packing.py): redshift the x51-oversampled band master curves onto the 300-bin model
// grid by linear interpolation, normalise over the bins, apply Fitzpatrick99 Milky Way dust at
// R_V = 3.1 and 1/(1+z), fold the band / distance scale, round to float32 and store only the
// non-zero extent of every band the SN was observed in.  Float64 arithmetic in the same order as
// the host path, so the float32 tiles agree with it to the last bit but for libm rounding.
//
// One warp per (SN, band); lanes over the 300 bins (10 per lane).  Two launches: supports
// (+ total width per SN, which fixes the tile stride), then the fill.
#pragma once
#include "bayesn_device.cuh"

namespace bayesn {

struct TilesDev {
    int N, n_b, n_bw;
    const double* master;      // [n_b][n_bw]
    const double* z;           // [N]
    const double* ebv;         // [N]
    const double* fold;        // [N]  log10 scale of SN s is bscale[k] - 0.4 * fold[s]
    const double* bscale;      // [n_b]
    const unsigned* used;      // [N] bit k set: band k has a valid observation
    const double* model_wave;  // [300]
    double band_spacing, RV_MW;
    double xk[9], yk[9], mk[9];   // F99 spline at R_V,MW: knots in 1/micron, knot values k(lambda-V), second derivatives
};

// natural cubic spline through (xk, yk) with second derivatives mk, linear extrapolation beyond the
// end knots (static.spline_matrix); FM90 ultraviolet branch below 2700 A (static.f99_alav)
__device__ __forceinline__ double f99_alav_dev(const TilesDev& T, double wave) {
    const double x = 1e4 / wave;
    double k;
    if (wave < 2700.0) {
        const double c2 = -0.824 + 4.717 / T.RV_MW, c1 = 2.030 - 3.007 * c2;
        const double x2 = x * x, y = x2 - 4.596 * 4.596;
        k = c1 + c2 * x + 3.23 * (x2 / (y * y + x2 * 0.99 * 0.99));
        if (x >= 5.9) {
            const double yy = x - 5.9;
            k += 0.41 * (0.5392 * yy * yy + 0.05644 * yy * yy * yy);
        }
    } else if (x > T.xk[8]) {
        const double h = T.xk[8] - T.xk[7], a = (T.xk[8] - x) / h;
        k = a * T.yk[7] + (1.0 - a) * T.yk[8] + (x - T.xk[8]) * h / 6.0 * T.mk[7];
    } else if (x < T.xk[0]) {
        const double h = T.xk[1] - T.xk[0], b = (x - T.xk[0]) / h;
        k = (1.0 - b) * T.yk[0] + b * T.yk[1] - (x - T.xk[0]) * h / 6.0 * T.mk[1];
    } else {
        int q = 0;
#pragma unroll
        for (int j = 1; j < 8; ++j)
            if (T.xk[j] <= x) q = j;
        const double h = T.xk[q + 1] - T.xk[q], a = (T.xk[q + 1] - x) / h, b = 1.0 - a;
        k = a * T.yk[q] + b * T.yk[q + 1] + ((a * a * a - a) * T.mk[q] + (b * b * b - b) * T.mk[q + 1]) * h * h / 6.0;
    }
    return 1.0 + k / T.RV_MW;
}

// the float32 row of (SN s, band k): lane holds bins lane + 32 i; returns the extent of non-zeros
__device__ __forceinline__ void tile_row(const TilesDev& T, int s, int k, int lane, float (&w32)[10], int& first,
                                         int& last) {
    const double z = T.z[s], opz = 1.0 + z;
    const double shift = log10(opz) / T.band_spacing;
    const double* mrow = T.master + (size_t)k * T.n_bw;
    double w[10], sum = 0.0;
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const int b = lane + 32 * i;
        w[i] = 0.0;
        if (b < NB) {
            const double loc = (double)(b * 51) + shift;
            const int il = (int)loc;
            const double rem = loc - (double)il;
            w[i] = rem * mrow[il + 1] + (1.0 - rem) * mrow[il];
            sum += w[i];
        }
    }
    for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(FULL, sum, o);
    const double scale = pow(10.0, T.bscale[k] - 0.4 * T.fold[s]);
    const double rve = T.RV_MW * T.ebv[s];
    first = NB;
    last = 0;
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const int b = lane + 32 * i;
        float v = 0.f;
        if (b < NB) {
            double x = w[i] / sum;
            const double mw = f99_alav_dev(T, T.model_wave[b] * opz) * rve;
            x = x * pow(10.0, -0.4 * mw);
            x = x / opz;
            v = (float)(x * scale);
            if (v != 0.f) { first = min(first, b); last = max(last, b + 1); }
        }
        w32[i] = v;
    }
    for (int o = 16; o; o >>= 1) {
        first = min(first, __shfl_xor_sync(FULL, first, o));
        last = max(last, __shfl_xor_sync(FULL, last, o));
    }
    if (last == 0) first = 0;
}

// (1) supports: support[s][k] = {first, last, 0, 0} (offset filled by the host-side scan kernel below)
__global__ void __launch_bounds__(128) tiles_support_kernel(TilesDev T, int4* __restrict__ support) {
    const int lane = threadIdx.x & 31;
    const long long unit = (long long)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (unit >= (long long)T.N * T.n_b) return;
    const int s = (int)(unit / T.n_b), k = (int)(unit - (long long)s * T.n_b);
    int first = 0, last = 0;
    if ((T.used[s] >> k) & 1u) {
        float w32[10];
        tile_row(T, s, k, lane, w32, first, last);
    }
    if (lane == 0) support[unit] = make_int4(first, last, 0, 0);
}

// (2) offsets of the rows inside each SN's tile + total width per SN
__global__ void tiles_offsets_kernel(int N, int n_b, int4* __restrict__ support, int* __restrict__ total) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= N) return;
    int off = 0;
    for (int k = 0; k < n_b; ++k) {
        int4 sp = support[(size_t)s * n_b + k];
        sp.z = off;
        off += sp.y - sp.x;
        support[(size_t)s * n_b + k] = sp;
    }
    total[s] = off;
}

// (3) fill: recompute the row and store its non-zero extent at its offset
__global__ void __launch_bounds__(128) tiles_fill_kernel(TilesDev T, const int4* __restrict__ support, int wt_stride,
                                                         float* __restrict__ weights) {
    const int lane = threadIdx.x & 31;
    const long long unit = (long long)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (unit >= (long long)T.N * T.n_b) return;
    const int s = (int)(unit / T.n_b), k = (int)(unit - (long long)s * T.n_b);
    const int4 sp = support[unit];
    if (sp.y <= sp.x) return;
    float w32[10];
    int first, last;
    tile_row(T, s, k, lane, w32, first, last);
    float* row = weights + (size_t)s * wt_stride + sp.z;
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const int b = lane + 32 * i;
        if (b >= sp.x && b < sp.y) row[b - sp.x] = w32[i];
    }
}

}  // namespace bayesn

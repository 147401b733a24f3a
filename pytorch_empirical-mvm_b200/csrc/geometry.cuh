// Closed-form shifted-window geometry shared by host and device.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define VSW_HD __host__ __device__ __forceinline__
#else
#define VSW_HD inline
#endif

namespace vsw {

struct WinGeom {
    int D, H, W;        // unpadded token grid
    int Dp, Hp, Wp;     // padded to multiples of the effective window (video_swin.py:213-218)
    int wd, wh, ww;     // effective window (after get_window_size, video_swin.py:95-108)
    int sd, sh, sw;     // effective shift
    int nWh, nWw;       // windows along h, w
    int nW, N;          // windows per clip, tokens per window

    // Token read by window `win`, slot `n` after roll(-shift) + window_partition
    // (video_swin.py:84-88, 220-229).  Returns the flat index in the UNPADDED grid, or -1 when the
    // source lies in the zero padding appended after norm1 (video_swin.py:217).
    VSW_HD int source_token(int win, int n) const {
        const int c = win % nWw, b = (win / nWw) % nWh, a = win / (nWw * nWh);
        const int k = n % ww, j = (n / ww) % wh, i = n / (ww * wh);
        int d = a * wd + i + sd; if (d >= Dp) d -= Dp;
        int h = b * wh + j + sh; if (h >= Hp) h -= Hp;
        int w = c * ww + k + sw; if (w >= Wp) w -= Wp;
        if (d >= D || h >= H || w >= W) return -1;
        return (d * H + h) * W + w;
    }

    // Region id of one axis position p in the SHIFTED frame (video_swin.py:296-298): the slices
    // [:-w], [-w:-s], [-s:] are assigned in that order, so s == 0 makes the last one cover the axis.
    static VSW_HD int axis_region(int p, int S, int w, int s) {
        if (s == 0) return 2;
        if (p < S - w) return 0;
        if (p < S - s) return 1;
        return 2;
    }

    // cnt = 9*rd + 3*rh + rw of slot n of window win (no roll: the counter image lives in the shifted frame)
    VSW_HD int region_id(int win, int n) const {
        const int c = win % nWw, b = (win / nWw) % nWh, a = win / (nWw * nWh);
        const int k = n % ww, j = (n / ww) % wh, i = n / (ww * wh);
        return 9 * axis_region(a * wd + i, Dp, wd, sd) + 3 * axis_region(b * wh + j, Hp, wh, sh) +
               axis_region(c * ww + k, Wp, ww, sw);
    }
};

inline bool make_win_geom(int D, int H, int W, int wd, int wh, int ww, int sd, int sh, int sw, WinGeom* g) {
    if (D <= 0 || H <= 0 || W <= 0 || wd <= 0 || wh <= 0 || ww <= 0) return false;
    if (sd < 0 || sh < 0 || sw < 0 || sd >= wd || sh >= wh || sw >= ww) return false;
    g->D = D; g->H = H; g->W = W;
    g->wd = wd; g->wh = wh; g->ww = ww;
    g->sd = sd; g->sh = sh; g->sw = sw;
    g->Dp = (D + wd - 1) / wd * wd;
    g->Hp = (H + wh - 1) / wh * wh;
    g->Wp = (W + ww - 1) / ww * ww;
    g->nWh = g->Hp / wh; g->nWw = g->Wp / ww;
    g->nW = (g->Dp / wd) * g->nWh * g->nWw;
    g->N = wd * wh * ww;
    return true;
}

}  // namespace vsw

/* eedi2_port.c -- TEST INFRASTRUCTURE (see oracle_port.h).
 *
 * Restates HandBrake's EEDI2 (libhb/templates/eedi2_template.c, driven by eedi2_interpolate_plane,
 * libhb/templates/decomb_template.c:366-441) for one field of a yuv420 frame, postproc 0..3.
 * Written as whole-image stage functions over a private copy of libhb's buffer layout:
 *   - every work buffer is one allocation holding the three planes back to back with libhb's
 *     64-byte stride (hb_frame_buffer_init), zero slack in front and behind, because several
 *     stages address x-1-u / x+1+u linearly and run into the neighbouring row or plane;
 *   - the edge-mask buffer persists between calls (only its top half is cleared per call, :132).
 * Stage order and buffer roles follow decomb.c:64-74 and decomb template :391-430.
 */
#include "oracle_port.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define LEAD 1024

typedef struct
{
    int depth, bps, peak, neutral, shift, shift2;
    int w[3], h[3], pitch[3], hh[3];
    size_t hoff[3], foff[3], hbytes, fbytes;
    uint8_t *half[4], *full[5];     /* SRCPF MSKPF TMPPF DSTPF ; DST2PF TMP2PF2 MSK2PF TMP2PF DST2MPF */
    int lim[33];
    int mthresh, vthresh, lthresh, dstr, estr, nt, maxd, pp;
    int *deriv[3][4];               /* postproc 2/3: x2, y2, xy, tmp per plane (see post_process_corner_stage) */
} eedi2_t;

static const int limlut_base[33] = { 6, 6, 7, 7, 8, 8, 9, 9, 9, 10, 10, 11, 11, 12, 12, 12, 12, 12, 12, 12,
                                     12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12, -1, -1 };

static inline int rd(const eedi2_t *e, const uint8_t *p, ptrdiff_t i)
{
    return e->bps == 2 ? ((const uint16_t *)p)[i] : p[i];
}
static inline void wr(const eedi2_t *e, uint8_t *p, ptrdiff_t i, int v)
{
    if (e->bps == 2) ((uint16_t *)p)[i] = (uint16_t)v;
    else p[i] = (uint8_t)v;
}
static inline int pixwrap(const eedi2_t *e, int v) { return e->bps == 2 ? (uint16_t)v : (uint8_t)v; }
static inline int iabs(int a) { return a < 0 ? -a : a; }
static inline int imax(int a, int b) { return a > b ? a : b; }
static inline int imin(int a, int b) { return a < b ? a : b; }

static void sort_metrics(int *order, int length)       /* eedi2.c:66-80 */
{
    for (int i = 1; i < length; ++i)
    {
        int j = i;
        const int temp = order[j];
        while (j > 0 && order[j - 1] > temp) { order[j] = order[j - 1]; --j; }
        order[j] = temp;
    }
}
static int median_of(const int *o, int n) { return (n & 1) ? o[n >> 1] : (o[(n - 1) >> 1] + o[n >> 1] + 1) >> 1; }
static int avg_round(int sum, int mid, int count) { return (int)(((float)(sum + mid) / (float)(count + 1)) + 0.5f); }

void *oracle_eedi2_create(int width, int height, int depth, int mthresh, int vthresh, int lthresh, int dstr, int estr,
                          int nt, int maxd, int pp)
{
    eedi2_t *e = calloc(1, sizeof(*e));
    e->depth = depth; e->bps = depth > 8 ? 2 : 1; e->peak = (1 << depth) - 1; e->neutral = 1 << (depth - 1);
    e->shift = depth - 8; e->shift2 = 2 + depth - 8;
    e->mthresh = mthresh; e->vthresh = vthresh; e->lthresh = lthresh; e->dstr = dstr; e->estr = estr;
    e->nt = nt; e->maxd = maxd; e->pp = pp;
    const int cw = -((-width) >> 1), ch = -((-height) >> 1);
    const int hfh = height / 2, hch = -((-hfh) >> 1);            /* decomb.c:291-296 */
    const int W[3] = { width, cw, cw }, H[3] = { height, ch, ch }, HH[3] = { hfh, hch, hch };
    for (int p = 0; p < 3; p++)
    {
        e->w[p] = W[p]; e->h[p] = H[p]; e->hh[p] = HH[p];
        e->pitch[p] = ((W[p] * e->bps + 63) / 64 * 64) / e->bps;
        e->hoff[p] = e->hbytes; e->foff[p] = e->fbytes;
        e->hbytes += (size_t)e->pitch[p] * HH[p] * e->bps;
        e->fbytes += (size_t)e->pitch[p] * H[p] * e->bps;
    }
    const size_t tail = (size_t)4 * e->pitch[0] * e->bps;
    for (int k = 0; k < 4; k++) e->half[k] = (uint8_t *)calloc(1, LEAD + e->hbytes + tail) + LEAD;
    for (int k = 0; k < 5; k++) e->full[k] = (uint8_t *)calloc(1, LEAD + e->fbytes + tail) + LEAD;
    if (pp > 1)
        for (int p = 0; p < 3; p++)
            for (int k = 0; k < 4; k++) e->deriv[p][k] = calloc((size_t)e->pitch[p] * (HH[p] + 1) + 16, sizeof(int));
    for (int i = 0; i < 33; i++)
        e->lim[i] = e->bps == 2 ? (uint16_t)(((uint16_t)limlut_base[i]) << e->shift) : (uint8_t)(((uint8_t)limlut_base[i]) << e->shift);
    return e;
}

void oracle_eedi2_destroy(void *ev)
{
    eedi2_t *e = ev;
    if (!e) return;
    for (int k = 0; k < 4; k++) free(e->half[k] - LEAD);
    for (int k = 0; k < 5; k++) free(e->full[k] - LEAD);
    for (int p = 0; p < 3; p++)
        for (int k = 0; k < 4; k++) free(e->deriv[p][k]);
    free(e);
}

/* ------------------------------------------------------------------ stages (one plane) */
static void blit(const eedi2_t *e, const uint8_t *src, uint8_t *dst, int pitch, int width, int rows)
{
    for (int y = 0; y < rows; y++) memcpy(dst + (size_t)y * pitch * e->bps, src + (size_t)y * pitch * e->bps, (size_t)width * e->bps);
}

static void edge_mask(const eedi2_t *e, uint8_t *dstp, const uint8_t *srcp, int pitch, int width, int height)
{
    /* callee parameter order (mthresh, lthresh, vthresh) vs call order (magnitude, variance, laplacian) */
    const int mthresh = e->mthresh * 10, lthresh = e->vthresh, vthresh = e->lthresh * 81;
    const int ten = pixwrap(e, 10 << e->shift), s = e->shift;
    memset(dstp, 0, (size_t)(height / 2) * pitch * e->bps);
    for (int y = 1; y < height - 1; ++y)
        for (int x = 1; x < width - 1; ++x)
        {
            const ptrdiff_t o = (ptrdiff_t)y * pitch + x;
            const int pm = rd(e, srcp, o - pitch - 1), pc = rd(e, srcp, o - pitch), pp = rd(e, srcp, o - pitch + 1);
            const int cm = rd(e, srcp, o - 1), cc = rd(e, srcp, o), cp = rd(e, srcp, o + 1);
            const int nm = rd(e, srcp, o + pitch - 1), nc = rd(e, srcp, o + pitch), np = rd(e, srcp, o + pitch + 1);
            if ((iabs(pc - cc) < ten && iabs(cc - nc) < ten && iabs(pc - nc) < ten) ||
                (iabs(pm - cm) < ten && iabs(cm - nm) < ten && iabs(pm - nm) < ten &&
                 iabs(pp - cp) < ten && iabs(cp - np) < ten && iabs(pp - np) < ten))
                continue;
            const int sum = (pm + pc + pp + cm + cc + cp + nm + nc + np) >> s;
            const int sumsq = (pm >> s) * (pm >> s) + (pc >> s) * (pc >> s) + (pp >> s) * (pp >> s) + (cm >> s) * (cm >> s) +
                              (cc >> s) * (cc >> s) + (cp >> s) * (cp >> s) + (nm >> s) * (nm >> s) + (nc >> s) * (nc >> s) + (np >> s) * (np >> s);
            if (9 * sumsq - sum * sum < vthresh) continue;
            const int Ix = (cp - cm) >> s;
            const int Iy = imax(imax(iabs(pc - nc), iabs(pc - cc)), iabs(cc - nc)) >> s;
            if (Ix * Ix + Iy * Iy >= mthresh) { wr(e, dstp, o, e->peak); continue; }
            const int Ixx = (cm - 2 * cc + cp) >> s, Iyy = (pc - 2 * cc + nc) >> s;
            if (iabs(Ixx) + iabs(Iyy) >= lthresh) wr(e, dstp, o, e->peak);
        }
}

static void morph(const eedi2_t *e, const uint8_t *mskp, uint8_t *dstp, int pitch, int width, int height, int str, int dilate)
{
    blit(e, mskp, dstp, pitch, width, height);
    for (int y = 1; y < height - 1; ++y)
        for (int x = 1; x < width - 1; ++x)
        {
            const ptrdiff_t o = (ptrdiff_t)y * pitch + x;
            const int v = rd(e, mskp, o);
            if (dilate ? v != 0 : v != e->peak) continue;
            int count = 0;
            for (int dy = -1; dy <= 1; dy++)
                for (int dx = -1; dx <= 1; dx++)
                    if ((dy || dx) && rd(e, mskp, o + (ptrdiff_t)dy * pitch + dx) == e->peak) ++count;
            if (dilate) { if (count >= str) wr(e, dstp, o, e->peak); }
            else        { if (count < str) wr(e, dstp, o, 0); }
        }
}

static void small_gaps(const eedi2_t *e, const uint8_t *mskp, uint8_t *dstp, int pitch, int width, int height)
{
    blit(e, mskp, dstp, pitch, width, height);
    for (int y = 1; y < height - 1; ++y)
        for (int x = 3; x < width - 3; ++x)
        {
            const ptrdiff_t o = (ptrdiff_t)y * pitch + x;
#define M(k) rd(e, mskp, o + (k))
            if (M(0))
            {
                if (M(-3) || M(-2) || M(-1) || M(1) || M(2) || M(3)) continue;
                wr(e, dstp, o, 0);
            }
            else if ((M(1) && (M(-1) || M(-2) || M(-3))) || (M(2) && (M(-1) || M(-2))) || (M(3) && M(-1)))
                wr(e, dstp, o, e->peak);
#undef M
        }
}

static void fill_all(const eedi2_t *e, uint8_t *dstp, size_t n, int v)
{
    for (size_t i = 0; i < n; i++) wr(e, dstp, (ptrdiff_t)i, v);
}

static void calc_directions(const eedi2_t *e, int plane, const uint8_t *mskp, const uint8_t *srcp, uint8_t *dstp,
                            int pitch, int width, int height)
{
    const int nt13 = pixwrap(e, (e->nt << e->shift) * 13), nt19 = pixwrap(e, (e->nt << e->shift) * 19);
    const int maxdt = plane == 0 ? e->maxd : (e->maxd >> 1), peak = e->peak;
    fill_all(e, dstp, (size_t)pitch * height, peak);
    for (int y = 1; y < height - 1; ++y)
        for (int x = 1; x < width - 1; ++x)
        {
            const ptrdiff_t o = (ptrdiff_t)y * pitch + x;
#define S(r, k) rd(e, srcp, o + (ptrdiff_t)(r) * pitch + (k))
#define MK(r, k) rd(e, mskp, o + (ptrdiff_t)(r) * pitch + (k))
            if (MK(0, 0) != peak || (MK(0, -1) != peak && MK(0, 1) != peak)) continue;
            const int startu = imax(-x + 1, -maxdt), stopu = imin(width - 2 - x, maxdt);
            const int base = iabs(S(0, 0) - S(1, 0)) + iabs(S(0, 0) - S(-1, 0));
            int minb = imin(nt13, base * 6), mina = imin(nt19, base * 9), minc = mina, mind = minb, mine = minb;
            int dira = -5000, dirb = -5000, dirc = -5000, dird = -5000, dire = -5000;
            for (int u = startu; u <= stopu; ++u)
            {
                if (!(y == 1 || MK(-1, -1 + u) == peak || MK(-1, u) == peak || MK(-1, 1 + u) == peak)) continue;
                if (!(y == height - 2 || MK(1, -1 - u) == peak || MK(1, -u) == peak || MK(1, 1 - u) == peak)) continue;
                const int diffsn = iabs(S(0, -1) - S(1, -1 - u)) + iabs(S(0, 0) - S(1, -u)) + iabs(S(0, 1) - S(1, 1 - u));
                const int diffsp = iabs(S(0, -1) - S(-1, -1 + u)) + iabs(S(0, 0) - S(-1, u)) + iabs(S(0, 1) - S(-1, 1 + u));
                const int diffps = iabs(S(-1, -1) - S(0, -1 - u)) + iabs(S(-1, 0) - S(0, -u)) + iabs(S(-1, 1) - S(0, 1 - u));
                const int diffns = iabs(S(1, -1) - S(0, -1 + u)) + iabs(S(1, 0) - S(0, u)) + iabs(S(1, 1) - S(0, 1 + u));
                const int diff = diffsn + diffsp + diffps + diffns;
                int diffd = diffsp + diffns, diffe = diffsn + diffps;
                if (diff < minb) { dirb = u; minb = diff; }
                if (y > 1)
                {
                    const int diff2pp = iabs(S(-2, -1) - S(-1, -1 - u)) + iabs(S(-2, 0) - S(-1, -u)) + iabs(S(-2, 1) - S(-1, 1 - u));
                    const int diffp2p = iabs(S(-1, -1) - S(-2, -1 + u)) + iabs(S(-1, 0) - S(-2, u)) + iabs(S(-1, 1) - S(-2, 1 + u));
                    const int diffa = diff + diff2pp + diffp2p;
                    diffd += diffp2p; diffe += diff2pp;
                    if (diffa < mina) { dira = u; mina = diffa; }
                }
                if (y < height - 2)
                {
                    const int diff2nn = iabs(S(2, -1) - S(1, -1 + u)) + iabs(S(2, 0) - S(1, u)) + iabs(S(2, 1) - S(1, 1 + u));
                    const int diffn2n = iabs(S(1, -1) - S(2, -1 - u)) + iabs(S(1, 0) - S(2, -u)) + iabs(S(1, 1) - S(2, 1 - u));
                    const int diffc = diff + diff2nn + diffn2n;
                    diffd += diff2nn; diffe += diffn2n;
                    if (diffc < minc) { dirc = u; minc = diffc; }
                }
                if (diffd < mind) { dird = u; mind = diffd; }
                if (diffe < mine) { dire = u; mine = diffe; }
            }
#undef S
#undef MK
            int order[5], k = 0, out = e->neutral;
            if (dira != -5000) order[k++] = dira;
            if (dirb != -5000) order[k++] = dirb;
            if (dirc != -5000) order[k++] = dirc;
            if (dird != -5000) order[k++] = dird;
            if (dire != -5000) order[k++] = dire;
            if (k > 1)
            {
                sort_metrics(order, k);
                const int mid = median_of(order, k);
                const int tlim = imax(e->lim[iabs(mid)] >> 2, 2);
                int sum = 0, count = 0;
                for (int i = 0; i < k; ++i)
                    if (iabs(order[i] - mid) <= tlim) { ++count; sum += order[i]; }
                if (count > 1) out = e->neutral + ((int)((float)sum / (float)count)) * (1 << e->shift2);
            }
            wr(e, dstp, o, out);
        }
}

/* filter_dir_map / expand_dir_map and their 2x variants */
static void dir_map(const eedi2_t *e, const uint8_t *mskp, const uint8_t *dmskp, uint8_t *dstp, int pitch, int width, int height,
                    int expand, int twox, int field)
{
    const int peak = e->peak;
    blit(e, dmskp, dstp, pitch, width, height);
    const int y0 = twox ? 2 - field : 1, ystep = twox ? 2 : 1;
    for (int y = y0; y < height - 1; y += ystep)
        for (int x = 1; x < width - 1; ++x)
        {
            const ptrdiff_t o = (ptrdiff_t)y * pitch + x;
            int active;
            if (twox) active = !(rd(e, mskp, o - pitch) != peak && rd(e, mskp, o + pitch) != peak);
            else      active = rd(e, mskp, o) == peak;
            if (expand) active = active && rd(e, dmskp, o) == peak;
            if (!active) continue;
            const ptrdiff_t st = twox ? 2 * (ptrdiff_t)pitch : pitch;
            const int up = !twox || y > 1, down = !twox || y < height - 2;
            int u = 0, order[9];
#define D(r, k) rd(e, dmskp, o + (r) * st + (k))
            if (up)   for (int k = -1; k <= 1; k++) if (D(-1, k) != peak) order[u++] = D(-1, k);
            if (D(0, -1) != peak) order[u++] = D(0, -1);
            if (!expand && D(0, 0) != peak) order[u++] = D(0, 0);
            if (D(0, 1) != peak) order[u++] = D(0, 1);
            if (down) for (int k = -1; k <= 1; k++) if (D(1, k) != peak) order[u++] = D(1, k);
            if (expand)
            {
                if (u < 5) continue;
            }
            else if (u < 4)
            {
                wr(e, dstp, o, peak);
                continue;
            }
            sort_metrics(order, u);
            const int mid = median_of(order, u);
            const int l = e->lim[iabs(mid - e->neutral) >> e->shift2];
            int sum = 0, count = 0;
            for (int i = 0; i < u; ++i)
                if (iabs(order[i] - mid) <= l) { ++count; sum += order[i]; }
            if (expand)
            {
                if (count < 5) continue;
            }
            else if (count < 4 || (count < 5 && D(0, 0) == peak))
            {
                wr(e, dstp, o, peak);
                continue;
            }
#undef D
            wr(e, dstp, o, avg_round(sum, mid, count));
        }
}

static void filter_map(const eedi2_t *e, const uint8_t *mskp, const uint8_t *dmskp, uint8_t *dstp, int pitch, int width, int height)
{
    const int peak = e->peak, shift = e->shift2, twelve = 12 << shift;
    blit(e, dmskp, dstp, pitch, width, height);
    for (int y = 1; y < height - 1; ++y)
        for (int x = 1; x < width - 1; ++x)
        {
            const ptrdiff_t o = (ptrdiff_t)y * pitch + x;
            const int cur = rd(e, dmskp, o);
            if (cur == peak || rd(e, mskp, o) != peak) continue;
            int dir = (cur - e->neutral) >> 2;
            const int lm = imax(iabs(dir) * 2, twelve);
            dir >>= shift;
            int ict = 0, icb = 0;
#define DC(j) rd(e, dmskp, o + (j))
#define DP(j) rd(e, dmskp, o - pitch + (j))
#define DN(j) rd(e, dmskp, o + pitch + (j))
#define BADP(j) ((iabs(DP(j) - cur) > lm && DP(j) != peak) || (DC(j) == peak && DP(j) == peak) || (iabs(DC(j) - cur) > lm && DC(j) != peak))
#define BADN(j) ((iabs(DN(j) - cur) > lm && DN(j) != peak) || (DN(j) == peak && DC(j) == peak) || (iabs(DC(j) - cur) > lm && DC(j) != peak))
            if (dir < 0) { for (int j = imax(-x, dir); j <= 0; ++j) if (BADP(j)) { ict = 1; break; } }
            else         { const int dt = imin(width - x - 1, dir); for (int j = 0; j <= dt; ++j) if (BADP(j)) { ict = 1; break; } }
            if (!ict) continue;
            if (dir < 0) { const int dt = imin(width - x - 1, iabs(dir)); for (int j = 0; j <= dt; ++j) if (BADN(j)) { icb = 1; break; } }
            else         { for (int j = imax(-x, -dir); j <= 0; ++j) if (BADN(j)) { icb = 1; break; } }
            if (icb) wr(e, dstp, o, peak);
#undef DC
#undef DP
#undef DN
#undef BADP
#undef BADN
        }
}

static void upscale2(const eedi2_t *e, const uint8_t *src, uint8_t *dst, int pitch, int rows)
{
    const size_t rb = (size_t)pitch * e->bps;
    for (int y = 0; y < rows; y++)
    {
        memcpy(dst + (size_t)(2 * y) * rb, src + (size_t)y * rb, rb);
        memcpy(dst + (size_t)(2 * y + 1) * rb, src + (size_t)y * rb, rb);
    }
}

static void mark_directions_2x(const eedi2_t *e, const uint8_t *mskp, const uint8_t *dmskp, uint8_t *dstp, int pitch, int width, int height, int tff)
{
    const int peak = e->peak;
    fill_all(e, dstp, (size_t)pitch * height, peak);
    for (int y = 2 - tff; y < height - 1; y += 2)
        for (int x = 1; x < width - 1; ++x)
        {
            const ptrdiff_t o0 = (ptrdiff_t)(y - 1) * pitch + x, o1 = o0 + 2 * (ptrdiff_t)pitch;
            if (rd(e, mskp, o0) != peak && rd(e, mskp, o1) != peak) continue;
            int v = 0, order[6];
            for (int k = -1; k <= 1; k++) if (rd(e, dmskp, o0 + k) != peak) order[v++] = rd(e, dmskp, o0 + k);
            for (int k = -1; k <= 1; k++) if (rd(e, dmskp, o1 + k) != peak) order[v++] = rd(e, dmskp, o1 + k);
            if (v < 3) continue;
            sort_metrics(order, v);
            const int mid = median_of(order, v);
            const int l = e->lim[iabs(mid - e->neutral) >> e->shift2];
            int u = 0;
#define A(k) rd(e, dmskp, o0 + (k))
#define B(k) rd(e, dmskp, o1 + (k))
            if (iabs(A(-1) - B(-1)) <= l || A(-1) == peak || B(-1) == peak) ++u;
            if (iabs(A(0) - B(0)) <= l || A(0) == peak || B(0) == peak) ++u;
            if (iabs(A(1) - B(-1)) <= l || A(1) == peak || B(1) == peak) ++u;          /* sic, :835 */
#undef A
#undef B
            if (u < 2) continue;
            int count = 0, sum = 0;
            for (int i = 0; i < v; ++i)
                if (iabs(order[i] - mid) <= l) { ++count; sum += order[i]; }
            if (count < v - 2 || count < 2) continue;
            wr(e, dstp, (ptrdiff_t)y * pitch + x, avg_round(sum, mid, count));
        }
}

static void fill_gaps_2x(const eedi2_t *e, const uint8_t *mskp, const uint8_t *dmskp, uint8_t *dstp, int pitch, int width, int height, int field)
{
    const int peak = e->peak, eight = 8 << e->shift, twenty = 20 << e->shift, fiveHundred = 500 << e->shift;
    blit(e, dmskp, dstp, pitch, width, height);
    for (int y = 2 - field; y < height - 1; y += 2)
    {
        const ptrdiff_t dc = (ptrdiff_t)y * pitch, dp = dc - 2 * (ptrdiff_t)pitch, dn = dc + 2 * (ptrdiff_t)pitch;
        const ptrdiff_t mc = (ptrdiff_t)(y - 1) * pitch, mpp = mc - 2 * (ptrdiff_t)pitch, mn = mc + 2 * (ptrdiff_t)pitch, mnn = mn + 2 * (ptrdiff_t)pitch;
        for (int x = 1; x < width - 1; ++x)
        {
            if (rd(e, dmskp, dc + x) != peak || (rd(e, mskp, mc + x) != peak && rd(e, mskp, mn + x) != peak)) continue;
            int u = x - 1, back = fiveHundred, forward = -fiveHundred;
            while (u)
            {
                if (rd(e, dmskp, dc + u) != peak) { back = rd(e, dmskp, dc + u); break; }
                if (rd(e, mskp, mc + u) != peak && rd(e, mskp, mn + u) != peak) break;
                --u;
            }
            int v = x + 1;
            while (v < width)
            {
                if (rd(e, dmskp, dc + v) != peak) { forward = rd(e, dmskp, dc + v); break; }
                if (rd(e, mskp, mc + v) != peak && rd(e, mskp, mn + v) != peak) break;
                ++v;
            }
            int tc = 1, bc = 1, mint = fiveHundred, maxt = -twenty, minb = fiveHundred, maxb = -twenty;
            for (int j = u; j <= v; ++j)
            {
                if (tc)
                {
                    if (y <= 2 || rd(e, dmskp, dp + j) == peak || (rd(e, mskp, mpp + j) != peak && rd(e, mskp, mc + j) != peak)) { tc = 0; mint = maxt = twenty; }
                    else { const int t = rd(e, dmskp, dp + j); if (t < mint) mint = t; if (t > maxt) maxt = t; }
                }
                if (bc)
                {
                    if (y >= height - 3 || rd(e, dmskp, dn + j) == peak || (rd(e, mskp, mn + j) != peak && rd(e, mskp, mnn + j) != peak)) { bc = 0; minb = maxb = twenty; }
                    else { const int t = rd(e, dmskp, dn + j); if (t < minb) minb = t; if (t > maxb) maxb = t; }
                }
            }
            if (maxt == -twenty) maxt = mint = twenty;
            if (maxb == -twenty) maxb = minb = twenty;
            const int fb = imax(iabs(forward - e->neutral), iabs(back - e->neutral));
            const int thresh = imax(imax(fb >> 2, eight), imax(iabs(mint - maxt), iabs(minb - maxb)));
            const int flim = imin(fb >> e->shift2, 6);
            if (iabs(forward - back) <= thresh && (v - u - 1 <= flim || tc || bc))
            {
                const double step = (double)(forward - back) / (double)(v - u);
                for (int j = 0; j < v - u - 1; ++j) wr(e, dstp, dc + u + j + 1, back + (int)(j * step + 0.5));
            }
        }
    }
}

/* interpolate_lattice, literally serial: the direction map is rewritten in place (:1148-1335) */
static void interpolate_lattice(const eedi2_t *e, int plane, uint8_t *dmskp, uint8_t *dstbase, const uint8_t *omskbase,
                                int pitch, int width, int height, int field)
{
    const int peak = e->peak, neutral = e->neutral, s = e->shift, shift2 = e->shift2;
    const int three = pixwrap(e, 3 << s), nine = pixwrap(e, 9 << s);
    const int nt4 = pixwrap(e, (e->nt << s) * 4), nt7 = pixwrap(e, (e->nt << s) * 7), nt8 = pixwrap(e, (e->nt << s) * 8);
    if (field == 1) memcpy(dstbase + (size_t)(height - 1) * pitch * e->bps, dstbase + (size_t)(height - 2) * pitch * e->bps, (size_t)width * e->bps);
    else            memcpy(dstbase, dstbase + (size_t)pitch * e->bps, (size_t)width * e->bps);
    for (int y = 2 - field; y < height - 1; y += 2)
    {
        const ptrdiff_t dm = (ptrdiff_t)y * pitch, a = (ptrdiff_t)(y - 1) * pitch, n = (ptrdiff_t)y * pitch, b = (ptrdiff_t)(y + 1) * pitch;
#define DM(k) rd(e, dmskp, dm + (k))
#define TP(k) rd(e, dstbase, a + (k))
#define BT(k) rd(e, dstbase, b + (k))
#define OP(k) rd(e, omskbase, a + (k))
#define ON(k) rd(e, omskbase, b + (k))
        for (int x = 0; x < width; ++x)
        {
            int dir = DM(x);
            const int cur = dir;
            const int l = e->lim[iabs(dir - neutral) >> shift2];
            const int avg = (TP(x) + BT(x) + 1) >> 1;
            if (dir == peak || (iabs(cur - DM(x - 1)) > l && iabs(cur - DM(x + 1)) > l))
            {
                wr(e, dstbase, n + x, avg);
                if (dir != peak) wr(e, dmskp, dm + x, neutral);
                continue;
            }
            if (l < nine)
            {
                const int sum = (TP(x - 1) + TP(x) + TP(x + 1) + BT(x - 1) + BT(x) + BT(x + 1)) >> s;
                const int sumsq = (TP(x - 1) >> s) * (TP(x - 1) >> s) + (TP(x) >> s) * (TP(x) >> s) + (TP(x + 1) >> s) * (TP(x + 1) >> s) +
                                  (BT(x - 1) >> s) * (BT(x - 1) >> s) + (BT(x) >> s) * (BT(x) >> s) + (BT(x + 1) >> s) * (BT(x + 1) >> s);
                if (6 * sumsq - sum * sum < 576)
                {
                    wr(e, dstbase, n + x, avg);
                    wr(e, dmskp, dm + x, peak);
                    continue;
                }
            }
            if (x > 1 && x < width - 2 &&
                ((TP(x) < imax(TP(x - 2), TP(x - 1)) - three && TP(x) < imax(TP(x + 2), TP(x + 1)) - three &&
                  BT(x) < imax(BT(x - 2), BT(x - 1)) - three && BT(x) < imax(BT(x + 2), BT(x + 1)) - three) ||
                 (TP(x) > imin(TP(x - 2), TP(x - 1)) + three && TP(x) > imin(TP(x + 2), TP(x + 1)) + three &&
                  BT(x) > imin(BT(x - 2), BT(x - 1)) + three && BT(x) > imin(BT(x + 2), BT(x + 1)) + three)))
            {
                wr(e, dstbase, n + x, avg);
                wr(e, dmskp, dm + x, neutral);
                continue;
            }
            dir = (dir - neutral + (1 << (shift2 - 1))) >> shift2;
            int val = avg;
            const int startu = (dir - 2 < 0) ? imax(-x + 1, imax(dir - 2, -width + 2 + x)) : imin(x - 1, imin(dir - 2, width - 2 - x));
            const int stopu  = (dir + 2 < 0) ? imax(-x + 1, imax(dir + 2, -width + 2 + x)) : imin(x - 1, imin(dir + 2, width - 2 - x));
            int mn = nt8;
#define NEARP(i) (OP(i) != peak && iabs(OP(i) - cur) <= l)
#define NEARN(i) (ON(i) != peak && iabs(ON(i) - cur) <= l)
            for (int u = startu; u <= stopu; ++u)
            {
                const int diff = iabs(TP(x - 1) - BT(x - u - 1)) + iabs(TP(x) - BT(x - u)) + iabs(TP(x + 1) - BT(x - u + 1)) +
                                 iabs(BT(x - 1) - TP(x + u - 1)) + iabs(BT(x) - TP(x + u)) + iabs(BT(x + 1) - TP(x + u + 1));
                if (!(diff < mn && (NEARP(x - 1 + u) || NEARP(x + u) || NEARP(x + 1 + u)) && (NEARN(x - 1 - u) || NEARN(x - u) || NEARN(x + 1 - u))))
                    continue;
                const int h0 = u >> 1, h1 = (u + 1) >> 1;
                const int diff2 = iabs(TP(x + h0 - 1) - BT(x - h0 - 1)) + iabs(TP(x + h0) - BT(x - h0)) + iabs(TP(x + h0 + 1) - BT(x - h0 + 1));
                if (!(diff2 < nt4 &&
                      (((iabs(OP(x + h0) - ON(x - h0)) <= l || iabs(OP(x + h0) - ON(x - h1)) <= l) && OP(x + h0) != peak) ||
                       ((iabs(OP(x + h1) - ON(x - h0)) <= l || iabs(OP(x + h1) - ON(x - h1)) <= l) && OP(x + h1) != peak))))
                    continue;
                if ((iabs(cur - OP(x + h0)) <= l || iabs(cur - OP(x + h1)) <= l) && (iabs(cur - ON(x - h0)) <= l || iabs(cur - ON(x - h1)) <= l))
                {
                    val = (TP(x + h0) + TP(x + h1) + BT(x - h0) + BT(x - h1) + 2) >> 2;
                    mn = diff;
                    dir = u;
                }
            }
#undef NEARP
#undef NEARN
            if (mn != nt8)
            {
                wr(e, dstbase, n + x, val);
                wr(e, dmskp, dm + x, neutral + dir * (1 << shift2));
            }
            else
            {
                const int minm = imin(TP(x), BT(x)), maxm = imax(TP(x), BT(x));
                const int d = plane == 0 ? 4 : 2;
                const int su = imax(-x + 1, -d), eu = imin(width - 2 - x, d);
                mn = nt7;
                for (int u = su; u <= eu; ++u)
                {
                    const int h0 = u >> 1, h1 = (u + 1) >> 1;
                    const int p1 = TP(x + h0) + TP(x + h1), p2 = BT(x - h0) + BT(x - h1);
                    const int diff = iabs(TP(x - 1) - BT(x - u - 1)) + iabs(TP(x) - BT(x - u)) + iabs(TP(x + 1) - BT(x - u + 1)) +
                                     iabs(BT(x - 1) - TP(x + u - 1)) + iabs(BT(x) - TP(x + u)) + iabs(BT(x + 1) - TP(x + u + 1)) + iabs(p1 - p2);
                    if (diff < mn)
                    {
                        const int valt = (p1 + p2 + 2) >> 2;
                        if (valt >= minm && valt <= maxm) { val = valt; mn = diff; dir = u; }
                    }
                }
                wr(e, dstbase, n + x, val);
                if (mn == 7 * e->nt) wr(e, dmskp, dm + x, neutral);
                else wr(e, dmskp, dm + x, neutral + dir * (1 << shift2));
            }
        }
#undef DM
#undef TP
#undef BT
#undef OP
#undef ON
    }
}

static void post_process(const eedi2_t *e, const uint8_t *nmskp, const uint8_t *omskp, uint8_t *dstp, int pitch, int width, int height, int field)
{
    for (int y = 2 - field; y < height - 1; y += 2)
        for (int x = 0; x < width; ++x)
        {
            const ptrdiff_t o = (ptrdiff_t)y * pitch + x;
            const int nm = rd(e, nmskp, o), om = rd(e, omskp, o);
            const int l = e->lim[iabs(nm - e->neutral) >> e->shift2];
            if (iabs(nm - om) > l && om != e->peak && om != e->neutral)
                wr(e, dstp, o, (rd(e, dstp, o - pitch) + rd(e, dstp, o + pitch) + 1) >> 1);
        }
}

/* ------------------------------------------------------------------ postproc 2/3: junctions and corners
 * (eedi2 template :1391-1904, called from decomb template :431-440).
 *
 * The reference's two blurs are written out as one expression per edge case.  Read together they say: a symmetric
 * kernel whose tap at distance k that would fall outside [0,n) is replaced by its point reflection through the centre
 * sample (x-k <-> x+k, hence the doubled weights), with ONE exception: in the horizontal pass of gaussian_blur_sqrt2
 * at x = width-2 the distance-3 tap reads srcp[x+3] for both sides (:1625) -- one element past the row, i.e. stride
 * padding or the next row's second element.  The vertical pass of that blur divides by a further 4 (>> 18).
 *
 * The reference runs its three plane threads concurrently over ONE set of scratch arrays (decomb.c:396-403, template
 * :380-383: a data race) and the exception above reads elements nothing ever wrote (malloc memory).  The contract
 * restated here is the race-free reading: every plane owns its scratch arrays, zero-filled when the filter starts. */
static const int blur1_taps[4] = { 26152, 15862, 3539, 291 };
static const int blur_sqrt2_taps[5] = { 18508, 14415, 6809, 1951, 339 };

static int blur_sample(const int *p, ptrdiff_t step, int i, int n, const int *taps, int radius, int typo_at, int typo_tap)
{
    int acc = p[0] * taps[0] + 32768;
    for (int k = 1; k <= radius; k++)
    {
        int a = i - k >= 0 ? -k : k, b = i + k <= n - 1 ? k : -k;
        if (i == typo_at && k == typo_tap) a = b = k;
        acc += (p[a * step] + p[b * step]) * taps[k];
    }
    return acc;
}

/* int arrays, horizontal then vertical; src == dst allowed (tmp holds the intermediate) */
static void blur_int(const int *src, int *tmp, int *dst, int pitch, int height, int width, const int *taps, int radius,
                     int typo, int vshift)
{
    for (int y = 0; y < height; y++)
        for (int x = 0; x < width; x++)
            tmp[(size_t)y * pitch + x] = blur_sample(src + (size_t)y * pitch + x, 1, x, width, taps, radius,
                                                     typo ? width - 2 : -1, 3) >> 16;
    for (int y = 0; y < height; y++)
        for (int x = 0; x < width; x++)
            dst[(size_t)y * pitch + x] = blur_sample(tmp + (size_t)y * pitch + x, pitch, y, height, taps, radius, -1, 0) >> vshift;
}

void oracle_eedi2_gaussian_blur_sqrt2(const int *src, int *tmp, int *dst, int pitch, int height, int width)
{
    blur_int(src, tmp, dst, pitch, height, width, blur_sqrt2_taps, 4, 1, 18);
}

/* pixel planes (:1402-1527); dst may be src */
void oracle_eedi2_gaussian_blur1(const void *src, void *tmp, void *dst, int pitch, int height, int width, int bps)
{
    eedi2_t e = { .bps = bps };
    int *a = malloc(((size_t)pitch * height + 8) * sizeof(int)), *b = malloc(((size_t)pitch * height + 8) * sizeof(int));
    for (size_t i = 0; i < (size_t)pitch * height; i++) a[i] = rd(&e, src, (ptrdiff_t)i);
    for (int y = 0; y < height; y++)
        for (int x = 0; x < width; x++)
        {
            b[(size_t)y * pitch + x] = pixwrap(&e, blur_sample(a + (size_t)y * pitch + x, 1, x, width, blur1_taps, 3, -1, 0) >> 16);
            wr(&e, tmp, (ptrdiff_t)y * pitch + x, b[(size_t)y * pitch + x]);
        }
    for (int y = 0; y < height; y++)
        for (int x = 0; x < width; x++)
            wr(&e, dst, (ptrdiff_t)y * pitch + x, blur_sample(b + (size_t)y * pitch + x, pitch, y, height, blur1_taps, 3, -1, 0) >> 16);
    free(a); free(b);
}

/* central differences, one-sided at the plane's edges, scaled back to 8 bits (:1756-1845) */
void oracle_eedi2_calc_derivatives(const void *src, int pitch, int height, int width, int *x2, int *y2, int *xy, int depth)
{
    eedi2_t e = { .bps = depth > 8 ? 2 : 1 };
    const int shift = depth - 8;
    for (int y = 0; y < height; y++)
        for (int x = 0; x < width; x++)
        {
            const ptrdiff_t o = (ptrdiff_t)y * pitch;
            const int Ix = (rd(&e, src, o + imin(x + 1, width - 1)) - rd(&e, src, o + imax(x - 1, 0))) >> shift;
            const int Iy = (rd(&e, src, (ptrdiff_t)imax(y - 1, 0) * pitch + x) - rd(&e, src, (ptrdiff_t)imin(y + 1, height - 1) * pitch + x)) >> shift;
            x2[o + x] = (Ix * Ix) >> 1;
            y2[o + x] = (Iy * Iy) >> 1;
            xy[o + x] = (Ix * Iy) >> 1;
        }
}

static int corner_response(int a, int b, int c)        /* Harris-style test value (:1882-1885) */
{
    return (int)(a * b - c * c - 0.09 * (a + b) * (a + b));
}

/* full-height plane rows y = 8-field, +2, ... < height-7 against derivative rows 3, 4, ... (:1864-1900) */
void oracle_eedi2_post_process_corner(const int *x2, const int *y2, const int *xy, int pitch, const void *mskp, void *dstp,
                                      int height, int width, int field, int depth)
{
    eedi2_t e = { .bps = depth > 8 ? 2 : 1 };
    const int neutral = 1 << (depth - 1), peak = (1 << depth) - 1;
    for (int y = 8 - field, r = 3; y < height - 7; y += 2, r++)
        for (int x = 4; x < width - 4; x++)
        {
            const ptrdiff_t o = (ptrdiff_t)y * pitch + x;
            const int m = rd(&e, mskp, o);
            if (m == peak || m == neutral) continue;
            const size_t d = (size_t)r * pitch + x;
            if (corner_response(x2[d], y2[d], xy[d]) > 775 || corner_response(x2[d + pitch], y2[d + pitch], xy[d + pitch]) > 775)
                wr(&e, dstp, o, (rd(&e, dstp, o - pitch) + rd(&e, dstp, o + pitch) + 1) >> 1);
        }
}

/* one field: `cur` = packed planar frame (tight rows); tff = pv->tff (decomb.c:539-542).  Result (full frame,
 * tight rows) into `out`.  The handle carries the edge-mask state from call to call. */
void oracle_eedi2_field(void *ev, const uint8_t *cur, int tff, uint8_t *out)
{
    eedi2_t *e = ev;
    size_t in_off = 0;
    for (int pl = 0; pl < 3; pl++)
    {
        const int pitch = e->pitch[pl], width = e->w[pl], height = e->h[pl], hh = e->hh[pl];
        uint8_t *srcp = e->half[0] + e->hoff[pl], *mskp = e->half[1] + e->hoff[pl], *tmpp = e->half[2] + e->hoff[pl], *dstp = e->half[3] + e->hoff[pl];
        uint8_t *dst2p = e->full[0] + e->foff[pl], *tmp2p2 = e->full[1] + e->foff[pl], *msk2p = e->full[2] + e->foff[pl],
                *tmp2p = e->full[3] + e->foff[pl], *dst2mp = e->full[4] + e->foff[pl];
        /* eedi2_planer: the field's lines with their whole stride; the harness frame has zero stride padding */
        for (int r = 0; r < (height + 1) / 2; r++)
        {
            uint8_t *d = srcp + (size_t)r * pitch * e->bps;
            memset(d, 0, (size_t)pitch * e->bps);
            memcpy(d, cur + in_off + (size_t)(2 * r + !tff) * width * e->bps, (size_t)width * e->bps);
        }
        edge_mask(e, mskp, srcp, pitch, width, hh);
        morph(e, mskp, tmpp, pitch, width, hh, e->estr, 0);
        morph(e, tmpp, mskp, pitch, width, hh, e->dstr, 1);
        morph(e, mskp, tmpp, pitch, width, hh, e->estr, 0);
        small_gaps(e, tmpp, mskp, pitch, width, hh);
        calc_directions(e, pl, mskp, srcp, tmpp, pitch, width, hh);
        dir_map(e, mskp, tmpp, dstp, pitch, width, hh, 0, 0, 0);
        dir_map(e, mskp, dstp, tmpp, pitch, width, hh, 1, 0, 0);
        filter_map(e, mskp, tmpp, dstp, pitch, width, hh);
        upscale2(e, srcp, dst2p, pitch, hh);
        upscale2(e, dstp, tmp2p2, pitch, hh);
        upscale2(e, mskp, msk2p, pitch, hh);
        mark_directions_2x(e, msk2p, tmp2p2, tmp2p, pitch, width, height, tff);
        dir_map(e, msk2p, tmp2p, dst2mp, pitch, width, height, 0, 1, tff);
        dir_map(e, msk2p, dst2mp, tmp2p, pitch, width, height, 1, 1, tff);
        fill_gaps_2x(e, msk2p, tmp2p, dst2mp, pitch, width, height, tff);
        fill_gaps_2x(e, msk2p, dst2mp, tmp2p, pitch, width, height, tff);
        interpolate_lattice(e, pl, tmp2p, dst2p, tmp2p2, pitch, width, height, tff);
        if (e->pp == 1 || e->pp == 3)
        {
            blit(e, tmp2p, tmp2p2, pitch, width, height);
            dir_map(e, msk2p, tmp2p, dst2mp, pitch, width, height, 0, 1, tff);
            dir_map(e, msk2p, dst2mp, tmp2p, pitch, width, height, 1, 1, tff);
            post_process(e, tmp2p, tmp2p2, dst2p, pitch, width, height, tff);
        }
        if (e->pp == 2 || e->pp == 3)
        {
            int *cx2 = e->deriv[pl][0], *cy2 = e->deriv[pl][1], *cxy = e->deriv[pl][2], *tmpc = e->deriv[pl][3];
            oracle_eedi2_gaussian_blur1(srcp, tmpp, srcp, pitch, hh, width, e->bps);
            oracle_eedi2_calc_derivatives(srcp, pitch, hh, width, cx2, cy2, cxy, e->depth);
            oracle_eedi2_gaussian_blur_sqrt2(cx2, tmpc, cx2, pitch, hh, width);
            oracle_eedi2_gaussian_blur_sqrt2(cy2, tmpc, cy2, pitch, hh, width);
            oracle_eedi2_gaussian_blur_sqrt2(cxy, tmpc, cxy, pitch, hh, width);
            oracle_eedi2_post_process_corner(cx2, cy2, cxy, pitch, tmp2p2, dst2p, height, width, tff, e->depth);
        }
        for (int y = 0; y < height; y++)
            memcpy(out + in_off + (size_t)y * width * e->bps, dst2p + (size_t)y * pitch * e->bps, (size_t)width * e->bps);
        in_off += (size_t)width * height * e->bps;
    }
}

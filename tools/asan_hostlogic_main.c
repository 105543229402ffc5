#include "hb_harness.h"
#include <stdio.h>
#include <stdlib.h>
extern hb_filter_object_t hb_filter_nlmeans_cuda, hb_filter_comb_detect_cuda, hb_filter_decomb_cuda, hb_filter_lapsharp_cuda,
                          hb_filter_unsharp_cuda, hb_filter_chroma_smooth_cuda, hb_filter_denoise_cuda, hb_filter_detelecine_cuda;
static unsigned rng = 777;
static unsigned rnd(void) { rng ^= rng << 13; rng ^= rng >> 17; rng ^= rng << 5; return rng; }
int main(void)
{
    struct { hb_filter_object_t *f; const char *s[3]; } jobs[] = {
        { &hb_filter_nlmeans_cuda, { NULL, "y-strength=6:y-frame-count=3:cb-range=5", "y-strength=4:y-prefilter=1032:threads=1" } },
        /* several device handles: per-device submission threads, block-cyclic dealing, halo copies (round 2) */
        { &hb_filter_nlmeans_cuda, { "y-strength=6:y-patch-size=3:devices=0,0:block=2", "y-strength=6:y-patch-size=3:y-frame-count=3:devices=0,0,0:block=2",
                                     "y-strength=4:y-patch-size=3:y-prefilter=1:devices=0,0,0,0:block=3:threads=1" } },
        { &hb_filter_lapsharp_cuda, { "y-strength=0.4:devices=0,0", "y-strength=0.4:devices=0,0,0", "y-strength=0.4:devices=0,0,0,0,0" } },
        { &hb_filter_comb_detect_cuda, { NULL, "mode=0:spatial-metric=0", "mode=2:spatial-metric=1:filter-mode=1" } },
        { &hb_filter_decomb_cuda, { "mode=7", "mode=31", "mode=55:parity=1" } },
        { &hb_filter_lapsharp_cuda, { NULL, "y-strength=1.1:y-kernel=lap:cb-strength=0.5:cb-kernel=isolog", "y-strength=0" } },
        { &hb_filter_unsharp_cuda, { NULL, "y-strength=1.5:y-size=15", "y-strength=0:cb-size=4" } },
        { &hb_filter_chroma_smooth_cuda, { NULL, "cb-strength=3:cb-size=15", "cb-strength=0" } },
        { &hb_filter_denoise_cuda, { NULL, "y-spatial=0:y-temporal=4", "y-spatial=300" } },
        { &hb_filter_detelecine_cuda, { NULL, "strict-breaks=1", "plane=2" } },
    };
    const int fmts[2] = { AV_PIX_FMT_YUV420P, AV_PIX_FMT_YUV420P10 };
    for (unsigned j = 0; j < sizeof(jobs) / sizeof(jobs[0]); j++)
        for (int k = 0; k < 3; k++)
            for (int d = 0; d < 2; d++)
            {
                const int fmt = fmts[d], w = 64 + 16 * (rnd() % 4), h = 48 + 8 * (rnd() % 4), n = 7;
                const size_t fb = hb_harness_frame_bytes(fmt, w, h);
                uint8_t *in = malloc(fb * n), *out = malloc(fb * (2 * n + 8));
                uint16_t flags[7]; uint8_t combed[7];
                for (int t = 0; t < n; t++)
                {
                    for (size_t i = 0; i < fb; i++) in[t * fb + i] = (uint8_t)((d && (i & 1)) ? rnd() % 4 : (i / 5 + t * 11 + rnd() % 40));
                    flags[t] = rnd() % 2 ? 0x0008 : 0x0010;
                    combed[t] = rnd() % 3;
                }
                int64_t start[40], stop[40]; double dur[40]; uint8_t comb[40]; uint16_t oflags[40];
                hb_harness_io_t io = { .pix_fmt = fmt, .width = w, .height = h, .n_in = n, .in = in, .in_flags = flags, .in_combed = combed, .out = out,
                                       .out_capacity = 2 * n + 8, .out_combed = comb, .out_flags = oflags, .out_start = start, .out_stop = stop, .out_duration = dur };
                if (hb_harness_run(jobs[j].f, jobs[j].s[k], &io) != 0) { printf("run failed: %s\n", jobs[j].f->name); return 1; }
                if (io.init_failed) { printf("init failed: %s %s\n", jobs[j].f->name, jobs[j].s[k] ? jobs[j].s[k] : ""); return 1; }
                free(in); free(out);
            }
    /* a device-resident chain (host memory stands in for device frames): reference counting of HBCU_DEVICE buffers */
    {
        extern hb_filter_object_t hb_filter_hbcu_upload, hb_filter_hbcu_download;
        extern long oracle_hbcu_frames_alive(void);
        hb_filter_object_t *chain[7] = { &hb_filter_hbcu_upload, &hb_filter_comb_detect_cuda, &hb_filter_decomb_cuda, &hb_filter_nlmeans_cuda,
                                         &hb_filter_lapsharp_cuda, &hb_filter_unsharp_cuda, &hb_filter_hbcu_download };
        const char *sets[7] = { NULL, NULL, "mode=63", "y-strength=6:y-patch-size=3", "y-strength=0.2", NULL, NULL };
        for (int d = 0; d < 2; d++)
        {
            const int fmt = fmts[d], w = 112, h = 64, n = 6;
            const size_t fb = hb_harness_frame_bytes(fmt, w, h);
            uint8_t *in = malloc(fb * n), *out = malloc(fb * (2 * n + 8));
            uint16_t flags[6];
            for (int t = 0; t < n; t++)
            {
                for (size_t i = 0; i < fb; i++) in[t * fb + i] = (uint8_t)((d && (i & 1)) ? rnd() % 4 : ((i / w) & 1 ? i / 3 + t * 29 : i / 3 + t * 5));
                flags[t] = 0x0008;
            }
            int64_t start[40], stop[40]; double dur[40]; uint8_t comb[40]; uint16_t oflags[40];
            hb_harness_io_t io = { .pix_fmt = fmt, .width = w, .height = h, .n_in = n, .in = in, .in_flags = flags, .out = out, .out_capacity = 2 * n + 8,
                                   .out_combed = comb, .out_flags = oflags, .out_start = start, .out_stop = stop, .out_duration = dur };
            if (hb_harness_run_chain(7, chain, sets, &io) != 0 || io.init_failed) { printf("device chain failed\n"); return 1; }
            free(in); free(out);
        }
        printf("device chain ran; device frames alive: %ld\n", oracle_hbcu_frames_alive());
    }
    printf("all filters ran; alive buffers: %ld\n", hb_shim_buffers_alive());
    return 0;
}

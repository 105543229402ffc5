/* oracle/ref_registry.c -- TEST INFRASTRUCTURE.
 * The part of libhb/common.c:5331-5537 (hb_filter_get / hb_filter_init /
 * hb_filter_close) that mt_frame_filter.c needs, restricted to the filters the
 * reference build in oracle/_ref contains. */
#include "handbrake/handbrake.h"

hb_filter_object_t *hb_filter_get(int filter_id)
{
    switch (filter_id)
    {
        case HB_FILTER_NLMEANS:     return &hb_filter_nlmeans;
        case HB_FILTER_COMB_DETECT: return &hb_filter_comb_detect;
        case HB_FILTER_DECOMB:      return &hb_filter_decomb;
        case HB_FILTER_LAPSHARP:    return &hb_filter_lapsharp;
        case HB_FILTER_DENOISE:     return &hb_filter_denoise;
        case HB_FILTER_UNSHARP:     return &hb_filter_unsharp;
        case HB_FILTER_CHROMA_SMOOTH: return &hb_filter_chroma_smooth;
        case HB_FILTER_DETELECINE:  return &hb_filter_detelecine;
        case HB_FILTER_MT_FRAME:    return &hb_filter_mt_frame;
        default:                    return NULL;
    }
}

/* common.c:5497-5517: lapsharp/unsharp/chroma_smooth are wrapped in mt_frame */
hb_filter_object_t *hb_filter_init(int filter_id)
{
    hb_filter_object_t *src = hb_filter_get(filter_id);
    if (src == NULL) return NULL;
    hb_filter_object_t *f = malloc(sizeof(*f));
    memcpy(f, src, sizeof(*f));
    if (filter_id == HB_FILTER_LAPSHARP || filter_id == HB_FILTER_UNSHARP || filter_id == HB_FILTER_CHROMA_SMOOTH)
    {
        hb_filter_object_t *wrapper = malloc(sizeof(*wrapper));
        memcpy(wrapper, &hb_filter_mt_frame, sizeof(*wrapper));
        wrapper->sub_filter = f;
        wrapper->id = filter_id;   /* wrapper takes the wrapped filter's id */
        return wrapper;
    }
    return f;
}

void hb_filter_close(hb_filter_object_t **pf)
{
    if (pf == NULL || *pf == NULL) return;
    hb_filter_object_t *f = *pf;
    if (f->sub_filter != NULL) hb_filter_close(&f->sub_filter);
    if (f->settings != NULL) hb_dict_free(&f->settings);
    free(f);
    *pf = NULL;
}

/* lapsharp as libhb instantiates it: wrapped in mt_frame (common.c:5497-5517).  A ready-made object
 * for the harness, which copies filter prototypes instead of calling hb_filter_init(). */
static hb_filter_object_t lapsharp_sub, unsharp_sub, chroma_smooth_sub;
hb_filter_object_t hb_filter_lapsharp_mt, hb_filter_unsharp_mt, hb_filter_chroma_smooth_mt;

static void wrap_mt(hb_filter_object_t *wrapper, hb_filter_object_t *sub, const hb_filter_object_t *src, int id)
{
    memcpy(sub, src, sizeof(*sub));
    memcpy(wrapper, &hb_filter_mt_frame, sizeof(*wrapper));
    wrapper->sub_filter = sub;
    wrapper->id = id;
}

__attribute__((constructor)) static void build_lapsharp_mt(void)
{
    memcpy(&lapsharp_sub, &hb_filter_lapsharp, sizeof(lapsharp_sub));
    memcpy(&hb_filter_lapsharp_mt, &hb_filter_mt_frame, sizeof(hb_filter_lapsharp_mt));
    hb_filter_lapsharp_mt.sub_filter = &lapsharp_sub;
    hb_filter_lapsharp_mt.id = HB_FILTER_LAPSHARP;
    wrap_mt(&hb_filter_unsharp_mt, &unsharp_sub, &hb_filter_unsharp, HB_FILTER_UNSHARP);
    wrap_mt(&hb_filter_chroma_smooth_mt, &chroma_smooth_sub, &hb_filter_chroma_smooth, HB_FILTER_CHROMA_SMOOTH);
}

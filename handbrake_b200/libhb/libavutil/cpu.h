/* libavutil/cpu.h -- shim: the hot path only asks FFmpeg for the SSE2 flag
 * (reference libhb/nlmeans_x86.c:153). */
#ifndef HBCU_SHIM_AVUTIL_CPU_H
#define HBCU_SHIM_AVUTIL_CPU_H
#define AV_CPU_FLAG_SSE2 0x0010
#ifdef __cplusplus
extern "C" {
#endif
int av_get_cpu_flags(void);
#ifdef __cplusplus
}
#endif
#endif

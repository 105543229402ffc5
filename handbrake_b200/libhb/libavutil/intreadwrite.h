/* libavutil/intreadwrite.h -- part of the shim: the two aligned native-endian 16-bit accessors libhb/denoise.c uses */
#ifndef HBCU_SHIM_INTREADWRITE_H
#define HBCU_SHIM_INTREADWRITE_H
#include <stdint.h>
#define AV_RN16A(p)    (*(const uint16_t *)(p))
#define AV_WN16A(p, v) (*(uint16_t *)(p) = (uint16_t)(v))
#endif

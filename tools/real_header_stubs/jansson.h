/* minimal jansson surface handbrake/hb_dict.h touches (types only; nothing is linked) */
#ifndef JANSSON_STUB_H
#define JANSSON_STUB_H
#include <stddef.h>
#include <stdio.h>
typedef struct json_t { int type; size_t refcount; } json_t;
typedef long long json_int_t;
typedef enum { JSON_OBJECT, JSON_ARRAY, JSON_STRING, JSON_INTEGER, JSON_REAL, JSON_TRUE, JSON_FALSE, JSON_NULL } json_type;
#endif

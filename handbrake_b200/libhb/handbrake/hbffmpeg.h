/* handbrake/hbffmpeg.h -- part of the shim; everything lives in handbrake.h */
#include "handbrake/handbrake.h"

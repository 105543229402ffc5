#include "av_stub.h"

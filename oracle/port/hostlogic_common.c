/* hostlogic_common.c -- TEST INFRASTRUCTURE: what the CPU stand-ins for the hbcu_* calls share (see hostlogic_*.c) */
#include <stdarg.h>
#include <stdio.h>

static char hostlogic_error[256] = "";

void oracle_hostlogic_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(hostlogic_error, sizeof(hostlogic_error), fmt, ap);
    va_end(ap);
}

const char *oracle_hbcu_last_error(void) { return hostlogic_error; }

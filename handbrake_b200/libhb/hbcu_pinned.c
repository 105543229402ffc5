/* hbcu_pinned.c -- "hb_buffer_t gains pinned backing": routes the frame-buffer
 * allocator (hb_buffer_init_internal, fifo.c:358-441) to page-locked memory
 * from the C-ABI so host<->device copies run asynchronously at full PCIe rate.
 * In libhb proper this is the one-line change in fifo.c shown in INTEGRATION.md. */
#include "handbrake/handbrake.h"
#include "hbcu.h"

void hbcu_use_pinned_buffers(int enable)
{
    if (enable) hb_shim_set_frame_allocator(hbcu_host_alloc, hbcu_host_free);
    else        hb_shim_set_frame_allocator(NULL, NULL);
}

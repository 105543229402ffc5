// hbcu_frames.h -- device frames (internal view).  The public surface is the hbcu_frame_* / hbcu_xfer_* part of
// include/hbcu.h; the filters use the begin/end helpers to order their streams against a frame's producer and readers.
#pragma once
#include "hbcu_common.h"

struct hbcu_frame_s
{
    int      device;
    uint8_t *base;                 // one allocation, planes back to back at `stride` (the layout of a STANDARD hb_buffer_t)
    size_t   bytes;
    uint8_t *plane[3];
    int      stride[3], row_bytes[3], rows[3];
    cudaEvent_t ready;             // recorded by the producer behind its last write
    cudaEvent_t consumed;          // recorded by every reader behind its last read; each reader first waits for the previous record,
                                   // so the latest record covers all readers
    int      refs;                 // hb_buffer_t references (hbcu_frame_retain / hbcu_frame_release)
    hbcu_frame_s *next;            // pool link
    // wrapped frames (hbcu_frame_wrap): memory owned by somebody else (a decoder surface); never pooled
    bool     external;
    void   (*ext_release)(void *);
    void    *ext_opaque;
};

namespace hbcu {

// producer side: `st` waits until every queued reader of the frame's previous life is done / marks the frame written
int frame_begin_write(hbcu_frame_s *f, cudaStream_t st);
int frame_end_write(hbcu_frame_s *f, cudaStream_t st);
// reader side: `st` waits for the producer / marks this reader done
int frame_begin_read(hbcu_frame_s *f, cudaStream_t st);
int frame_end_read(hbcu_frame_s *f, cudaStream_t st);

}  // namespace hbcu

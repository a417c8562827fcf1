/*
 * hlb200_glue.h -- batch mode of the drop-in glue (host/hlb200_glue.c): lets a driver that runs MANY codec instances (streams) encode one picture of each with ONE
 * device launch, through the reference's unchanged public API (hl_codec_encode, source/hl_codec.c:152).
 *
 * hl_codec_encode() runs the reference's host code up to the slice hook (SPS/PPS/slice header, DPB, reference lists), and the slice hook needs the device's answer
 * before it can return.  In batch mode the hook SUBMITS its picture and calls `yield`; the driver runs every stream's hl_codec_encode() that far (each on its own
 * stack: coroutines or threads), calls hlb200_glue_batch_flush() -- one hlb200_slice_encode_batch_async + one hlb200_slice_bits_batch_async for all of them -- and
 * lets the hooks continue: each downloads its slice data and finishes its NAL unit.  host/hl_b200_multi.c is such a driver (ucontext coroutines, one thread).
 */
#ifndef HLB200_GLUE_H_
#define HLB200_GLUE_H_
#ifdef __cplusplus
extern "C" {
#endif
typedef void (*hlb200_glue_yield_fn)(void* arg);
void hlb200_glue_batch_begin(hlb200_glue_yield_fn yield, void* arg);   /* from now on single-layer slice hooks submit and yield */
int hlb200_glue_batch_pending(void);                                    /* pictures submitted since the last flush */
int hlb200_glue_batch_flush(void);                                      /* encodes + serialises them with one launch each; HL_ERROR_T value */
void hlb200_glue_batch_end(void);
#ifdef __cplusplus
}
#endif
#endif

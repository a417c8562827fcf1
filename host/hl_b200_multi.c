/*
 * hl_b200_multi.c -- many H.264 streams through the reference's own public API (hl_engine_init / hl_codec_create / hl_codec_encode, include/hartallo/hl_api.h,
 * source/hl_codec.c:152) with the B200 hot path dropped in by host/hlb200_glue.c in BATCH mode: one device launch encodes one picture of every stream.
 *
 * It is the multi-stream counterpart of source/test_encoder.c:78-244 (same codec settings, :135-146; same fps print, :240-244, as JSON).  Every stream is an
 * unmodified hl_codec_t; hl_codec_encode() runs the reference's host code (SPS/PPS/slice header, DPB, POC, reference lists, NAL assembly, emulation prevention) and
 * the glue's slice hook submits the picture and yields.  The streams' encode calls run as coroutines (ucontext) on ONE thread -- the reference keeps process-wide
 * statics, so it is never entered concurrently -- in groups that alternate, so that the host work of one group overlaps the kernel of another:
 *
 *     submit A(f) | launch A(f) | finish B(f-1) (slice bits -> NAL) | submit B(f) | launch B(f) | finish A(f) | submit A(f+1) | launch A(f+1) | ...
 *
 * Output: one JSON line (streams, pictures, encode fps, macroblocks/s, bitstream bytes, MD5 of stream 0's bitstream; with --same-content every stream must produce
 * the bitstream of stream 0).  Host code stays C; the device is reached through libhl_b200.so's C-ABI only.
 */
#define _GNU_SOURCE
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <ucontext.h>
#include <sys/mman.h>

#include "hartallo/hl_api.h"
#include "hartallo/hl_codec.h"
#include "hartallo/hl_frame.h"
#include "hartallo/hl_debug.h"
#include "hartallo/hl_md5.h"

#include "hlb200.h"
#include "hlb200_glue.h"

enum { ST_IDLE = 0, ST_SUBMITTED, ST_FRAME_DONE, ST_FAILED };
typedef struct stream_s {
    struct hl_codec_s* codec;
    struct hl_codec_result_s* result;
    struct hl_frame_video_s* frame;
    ucontext_t uc;
    void* stack;
    int content, state, err, frames_done;
    uint8_t* out; size_t out_n, out_cap;
} stream_t;

static ucontext_t g_main;
static stream_t* g_running = NULL;
static int g_w = 1920, g_h = 1088, g_frames = 6, g_qp = 31, g_me_range = 32, g_gen = 1;
static uint8_t** g_content = NULL;   /* [distinct] -> frames x frame_bytes */
static size_t g_frame_bytes;

static double now_ms(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; }

/* G1 "pan" of SURVEY 8(d) with a per-sequence LCG (hartallo_b200/synth.py: G1) */
static void gen_g1(uint8_t* yuv, int w, int h, int n, uint32_t* lcg)
{
    int x, y, i;
    uint8_t* Y = yuv; uint8_t* UV = yuv + (size_t)w * h;
    for (y = 0; y < h; ++y) for (x = 0; x < w; ++x) {
        int v;
        *lcg = *lcg * 1664525u + 1013904223u;
        v = 128 + 60 * ((((x + 2 * n) / 8) + ((y + n) / 8)) & 1) + ((((x + 2 * n) * 7) + ((y + n) * 13)) & 31) - 16 + (int)((*lcg >> 8) & 3);
        Y[(size_t)y * w + x] = (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
    }
    for (i = 0; i < (w * h) / 2; ++i) UV[i] = (uint8_t)(128 + ((i + n) & 15));
}
/* G2 "stress" (hartallo_b200/synth.py: G2): random base picture with 8x8 patches of 0 / 255, translated by (3n, -2n) */
static void gen_g2(uint8_t* yuv, const uint8_t* base, int w, int h, int n)
{
    int x, y, c;
    for (y = 0; y < h; ++y) for (x = 0; x < w; ++x) yuv[(size_t)y * w + x] = base[(size_t)(((y - 2 * n) % h + h) % h) * w + (((x + 3 * n) % w + w) % w)];
    for (c = 0; c < 2; ++c) {
        const int cw = w / 2, ch = h / 2;
        const uint8_t* b = base + (size_t)w * h + (size_t)c * cw * ch;
        uint8_t* o = yuv + (size_t)w * h + (size_t)c * cw * ch;
        for (y = 0; y < ch; ++y) for (x = 0; x < cw; ++x) o[(size_t)y * cw + x] = b[(size_t)(((y - n) % ch + ch) % ch) * cw + (((x + (3 * n) / 2) % cw + cw) % cw)];
    }
}
static uint8_t* g2_base(int w, int h, uint32_t seed)
{
    uint32_t s = seed * 2654435761u + 97u;
    size_t i, tot = (size_t)w * h * 3 / 2;
    int x, y;
    uint8_t* b = (uint8_t*)malloc(tot);
    for (i = 0; i < tot; ++i) {
        uint32_t a, v;
        s = s * 1664525u + 1013904223u; a = s >> 8;
        s = s * 1664525u + 1013904223u; v = s >> 8;
        b[i] = (uint8_t)(v % ((a & 1) ? 34u : 256u));
    }
    for (y = 0; y + 8 <= h; y += 8) for (x = 0; x + 8 <= w; x += 8) {
        uint32_t a; int yy, xx;
        s = s * 1664525u + 1013904223u; a = (s >> 8) & 15;
        if (a < 2) for (yy = 0; yy < 8; ++yy) for (xx = 0; xx < 8; ++xx) b[(size_t)(y + yy) * w + x + xx] = (a == 0) ? 0 : 255;
    }
    return b;
}

static void yield_cb(void* arg)
{
    stream_t* s = g_running;
    (void)arg;
    s->state = ST_SUBMITTED;
    swapcontext(&s->uc, &g_main);
}

static void stream_main(unsigned lo, unsigned hi)
{
    stream_t* s = (stream_t*)(((uintptr_t)hi << 32) | (uintptr_t)lo);
    int i;
    for (i = 0; i < g_frames; ++i) {
        uint8_t* yuv = g_content[s->content] + (size_t)i * g_frame_bytes;
        HL_ERROR_T err = hl_frame_video_fill(s->frame, HL_VIDEO_CHROMA_YUV420, (uint32_t)g_w, (uint32_t)g_h, yuv, g_frame_bytes);
        s->frame->encoding = HL_VIDEO_ENCODING_TYPE_AUTO;
        if (!err) err = hl_codec_encode(s->codec, (hl_frame_t*)s->frame, s->result);
        if (err) { s->err = (int)err; s->state = ST_FAILED; swapcontext(&s->uc, &g_main); return; }
        {
            size_t need = s->out_n + 3 + s->result->data_size + (size_t)s->codec->hdr_bytes_count + 16;
            if (need > s->out_cap) { s->out_cap = need * 2; s->out = (uint8_t*)realloc(s->out, s->out_cap); }
        }
        if (s->result->type & HL_CODEC_RESULT_TYPE_HDR) { memcpy(s->out + s->out_n, s->codec->hdr_bytes, s->codec->hdr_bytes_count); s->out_n += s->codec->hdr_bytes_count; }
        if (s->result->type & HL_CODEC_RESULT_TYPE_DATA) {
            static const uint8_t scp[3] = { 0, 0, 1 };
            memcpy(s->out + s->out_n, scp, 3); s->out_n += 3;
            memcpy(s->out + s->out_n, s->result->data_ptr, s->result->data_size); s->out_n += s->result->data_size;
        }
        ++s->frames_done;
        s->state = ST_FRAME_DONE;
        swapcontext(&s->uc, &g_main);
    }
    s->state = ST_FRAME_DONE;
    for (;;) swapcontext(&s->uc, &g_main);
}

static int resume(stream_t* s)
{
    g_running = s;
    swapcontext(&g_main, &s->uc);
    g_running = NULL;
    if (s->state == ST_FAILED) { fprintf(stderr, "stream failed: HL_ERROR %d\n", s->err); return -1; }
    return 0;
}

static void md5_hex(const uint8_t* p, size_t n, char out[33])
{
    hl_md5context_t ctx; hl_md5digest_t d; int i;
    hl_md5init(&ctx); hl_md5update(&ctx, p, n); hl_md5final(d, &ctx);
    for (i = 0; i < 16; ++i) sprintf(out + 2 * i, "%02x", d[i]);
    out[32] = 0;
}

int main(int argc, char** argv)
{
    int streams = 8, distinct = 0, warm = 1, same = 0, groups = 2, refs = 1, deblock = 0, early = 0, i, f, g, failed = 0;
    uint32_t seed = 3;
    const char* out_path = NULL;
    const struct hl_codec_plugin_def_s* plugin = NULL;
    stream_t* st;
    double t0 = 0, t1, t_first = 0, t_submit = 0, t_finish = 0, t_flush = 0, t_finish_first = 0, ta;
    size_t bytes_timed = 0, bytes_at_t0 = 0;
    char md5[33];
    HL_ERROR_T err;
    for (i = 1; i < argc; ++i) {
        if (!strcmp(argv[i], "--size") && i + 2 < argc) { g_w = atoi(argv[++i]); g_h = atoi(argv[++i]); }
        else if (!strcmp(argv[i], "--streams") && i + 1 < argc) streams = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--frames") && i + 1 < argc) g_frames = atoi(argv[++i]);       /* per stream, IDR included */
        else if (!strcmp(argv[i], "--warmup") && i + 1 < argc) warm = atoi(argv[++i]);           /* untimed P pictures after the IDR */
        else if (!strcmp(argv[i], "--qp") && i + 1 < argc) g_qp = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--me-range") && i + 1 < argc) g_me_range = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--refs") && i + 1 < argc) refs = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--distinct") && i + 1 < argc) distinct = atoi(argv[++i]);    /* distinct contents (streams share them round robin); default: one per stream, at most 16 */
        else if (!strcmp(argv[i], "--same-content")) same = 1;                                   /* every stream encodes sequence 0: all bitstreams must be equal */
        else if (!strcmp(argv[i], "--groups") && i + 1 < argc) groups = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--seed") && i + 1 < argc) seed = (uint32_t)atoi(argv[++i]);
        else if (!strcmp(argv[i], "--out") && i + 1 < argc) out_path = argv[++i];               /* bitstream of stream 0 */
        else if (!strcmp(argv[i], "--gen") && i + 1 < argc) { ++i; g_gen = !strcmp(argv[i], "g2") ? 2 : 1; }
        else if (!strcmp(argv[i], "--deblock") && i + 1 < argc) deblock = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--early-term") && i + 1 < argc) early = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--defaults")) deblock = early = 1;                                /* what hl_codec_create leaves (hl_types.h:67,69) */
        else { fprintf(stderr, "unknown arg %s\n", argv[i]); return 2; }
    }
    if (streams < 1 || streams > 1024 || g_frames < 2 || (g_w & 15) || (g_h & 15)) { fprintf(stderr, "bad arguments\n"); return 2; }
    if (same) distinct = 1;
    if (distinct <= 0) distinct = streams < 16 ? streams : 16;
    if (groups < 1) groups = 1;
    if (groups > streams) groups = streams;
    if (warm > g_frames - 2) warm = g_frames - 2;
    g_frame_bytes = (size_t)g_w * g_h * 3 / 2;

    /* synthetic sequences in pinned host memory (what a capture pipeline hands an encoder) */
    g_content = (uint8_t**)calloc((size_t)distinct, sizeof(uint8_t*));
    for (i = 0; i < distinct; ++i) {
        uint32_t lcg = 12345u + 7919u * (uint32_t)i;     /* hartallo_b200/sharding.py: stream_seed */
        uint8_t* base = g_gen == 2 ? g2_base(g_w, g_h, seed + (uint32_t)i) : NULL;
        g_content[i] = (uint8_t*)malloc(g_frame_bytes * (size_t)g_frames);
        for (f = 0; f < g_frames; ++f) {
            if (g_gen == 1) gen_g1(g_content[i] + (size_t)f * g_frame_bytes, g_w, g_h, f, &lcg);
            else gen_g2(g_content[i] + (size_t)f * g_frame_bytes, base, g_w, g_h, f);
        }
        free(base);
        hlb200_host_register(g_content[i], g_frame_bytes * (size_t)g_frames);
    }

    hl_debug_set_level(HL_DEBUG_LEVEL_ERROR);
    hl_engine_set_cpu_flags(0);
    if ((err = hl_engine_init())) { fprintf(stderr, "engine init %d\n", err); return 1; }
    if ((err = hl_codec_plugin_find(HL_CODEC_TYPE_H264_SVC, &plugin))) { fprintf(stderr, "plugin find %d\n", err); return 1; }
    st = (stream_t*)calloc((size_t)streams, sizeof(stream_t));
    for (i = 0; i < streams; ++i) {
        stream_t* s = &st[i];
        if ((err = hl_codec_create(plugin, &s->codec)) || (err = hl_codec_result_create(&s->result)) || (err = hl_frame_video_create(&s->frame))) { fprintf(stderr, "create %d\n", err); return 1; }
        /* same knobs as source/test_encoder.c:135-146 */
        s->codec->gop_size = 400; s->codec->me_range = g_me_range; s->codec->qp = g_qp;
        s->codec->fps.num = 1; s->codec->fps.den = 30;
        s->codec->rc_bitrate = -1; s->codec->deblock_flag = deblock; s->codec->threads_count = 1; s->codec->max_ref_frame = refs;
        s->codec->distortion_mesure_type = HL_VIDEO_DISTORTION_MESURE_TYPE_SAD;
        s->codec->me_type = (HL_VIDEO_ME_TYPE_INTEGER | HL_VIDEO_ME_TYPE_HALF | HL_VIDEO_ME_TYPE_QUATER);
        s->codec->me_part_types = HL_VIDEO_ME_PART_TYPE_ALL; s->codec->me_subpart_types = HL_VIDEO_ME_SUBPART_TYPE_ALL;
        s->codec->me_early_term_flag = early;
        s->content = i % distinct;
        s->stack = mmap(NULL, 1u << 20, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_STACK, -1, 0);
        if (s->stack == MAP_FAILED) { perror("mmap"); return 1; }
        getcontext(&s->uc);
        s->uc.uc_stack.ss_sp = s->stack; s->uc.uc_stack.ss_size = 1u << 20; s->uc.uc_link = &g_main;
        makecontext(&s->uc, (void (*)(void))stream_main, 2, (unsigned)((uintptr_t)s & 0xffffffffu), (unsigned)((uintptr_t)s >> 32));
    }

    hlb200_glue_batch_begin(yield_cb, NULL);
    t_first = now_ms();
    /* software pipeline over the groups: while the device encodes picture f of one group, the host finishes picture f-1 of the next group (slice bits -> NAL unit)
     * and submits its picture f; the library runs batch launches back to back (hlb200_slice_encode_batch_async), so the device always has the next one queued */
    for (f = 0; f <= g_frames && !failed; ++f) {
        for (g = 0; g < groups && !failed; ++g) {
            const int a = (int)((long)streams * g / groups), b = (int)((long)streams * (g + 1) / groups);
            if (f == 1 + warm && g == 0) {   /* the timed region starts (and ends) with the pipeline drained: exactly `timed` pictures per stream lie inside it */
                for (i = 0; i < streams && !failed; ++i)
                    if (st[i].state == ST_SUBMITTED) failed |= resume(&st[i]) != 0;
                t0 = now_ms(); bytes_at_t0 = 0; t_submit = t_finish = t_flush = t_finish_first = 0;
                for (i = 0; i < streams; ++i) bytes_at_t0 += st[i].out_n;
            }
            if (f > 0) {   /* finish picture f-1 of the group: every hook downloads its slice data, the reference completes the NAL unit */
                ta = now_ms();
                for (i = a; i < b && !failed; ++i) {
                    const double tb = now_ms();
                    if (st[i].state == ST_SUBMITTED) failed |= resume(&st[i]) != 0;
                    if (i == a) t_finish_first += now_ms() - tb;   /* the first download of a group waits for the group's launch */
                }
                t_finish += now_ms() - ta;
                for (i = a; i < b && !failed; ++i)
                    if (st[i].state != ST_FRAME_DONE) { fprintf(stderr, "stream %d did not finish picture %d (state %d)\n", i, f - 1, st[i].state); failed = 1; }
            }
            if (f < g_frames) {   /* submit picture f of the group, one launch for all of them */
                ta = now_ms();
                for (i = a; i < b && !failed; ++i) failed |= resume(&st[i]) != 0;
                t_submit += now_ms() - ta; ta = now_ms();
                if (!failed && hlb200_glue_batch_pending() > 0 && (err = (HL_ERROR_T)hlb200_glue_batch_flush())) { fprintf(stderr, "batch launch failed: %d (%s)\n", (int)err, hlb200_last_error()); failed = 1; }
                t_flush += now_ms() - ta;
            }
        }
    }
    t1 = now_ms();
    hlb200_glue_batch_end();
    if (failed) return 1;
    for (i = 0; i < streams; ++i) bytes_timed += st[i].out_n;
    bytes_timed -= bytes_at_t0;
    md5_hex(st[0].out, st[0].out_n, md5);   /* the reference's own hl_md5 (as oracle/ref_driver.c prints it) */
    if (out_path) { FILE* fo = fopen(out_path, "wb"); if (!fo) { perror(out_path); return 2; } fwrite(st[0].out, 1, st[0].out_n, fo); fclose(fo); }
    {
        int all_equal = 1;
        const int timed = g_frames - 1 - warm, mbs = (g_w / 16) * (g_h / 16);
        const double ms = t1 - t0;
        if (same) for (i = 1; i < streams; ++i) all_equal &= st[i].out_n == st[0].out_n && !memcmp(st[i].out, st[0].out, st[0].out_n);
        printf("{\"streams\": %d, \"width\": %d, \"height\": %d, \"frames\": %d, \"timed_pictures_per_stream\": %d, \"groups\": %d, \"ms_timed\": %.3f, \"ms_total\": %.3f, "
               "\"encode_fps\": %.2f, \"fps_per_stream\": %.3f, \"mb_per_s\": %.1f, \"bitstream_bytes_timed\": %zu, \"bytes\": %zu, \"md5\": \"%s\", \"same_content\": %d, \"all_streams_equal\": %s, "
               "\"host_ms\": {\"submit\": %.1f, \"launch\": %.1f, \"finish\": %.1f, \"finish_first_of_group\": %.1f}}\n",
               streams, g_w, g_h, g_frames, timed, groups, ms, t1 - t_first, 1e3 * timed * streams / ms, 1e3 * timed / ms, 1e3 * (double)timed * streams * mbs / ms, bytes_timed,
               st[0].out_n, md5, same, same ? (all_equal ? "true" : "false") : "null", t_submit, t_flush, t_finish, t_finish_first);
        if (same && !all_equal) return 3;
    }
    return 0;
}

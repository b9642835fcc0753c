/* Test double of libmp3lame (tests only; LAME itself is not in the image): same entry points as the ones the
 * reference's export calls (src/processor/audio-io.cpp:656-822), but instead of encoding it RECORDS.  Every call is
 * appended as a text line to the file named by FAKE_LAME_LOG; every encode call "returns" one 16-byte record
 * {"FLAM", kind, nsamples, float checksum} as its MP3 bytes, so the written file shows call order and payload.
 * FAKE_LAME_FAIL=init_params | encode makes that call report an error. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct { FILE* log; int channels; } fake_t;

static void logf_(fake_t* f, const char* fmt, long a, long b)
{
    if (!f || !f->log) return;
    fprintf(f->log, fmt, a, b);
    fputc('\n', f->log);
    fflush(f->log);
}
static int failing(const char* what)
{
    const char* e = getenv("FAKE_LAME_FAIL");
    return e && !strcmp(e, what);
}

void* lame_init(void)
{
    fake_t* f = (fake_t*)calloc(1, sizeof(fake_t));
    const char* path = getenv("FAKE_LAME_LOG");
    if (path) f->log = fopen(path, "a");
    logf_(f, "init", 0, 0);
    return f;
}
int lame_close(void* h)
{
    fake_t* f = (fake_t*)h;
    logf_(f, "close", 0, 0);
    if (f->log) fclose(f->log);
    free(f);
    return 0;
}
#define SETTER(name)                                                    \
    int lame_set_##name(void* h, int v) { logf_((fake_t*)h, "set_" #name " %ld", v, 0); return 0; }
SETTER(in_samplerate)
SETTER(quality)
SETTER(mode)
SETTER(out_samplerate)
SETTER(VBR)
SETTER(brate)
int lame_set_num_channels(void* h, int v) { ((fake_t*)h)->channels = v; logf_((fake_t*)h, "set_num_channels %ld", v, 0); return 0; }
int lame_init_params(void* h)
{
    logf_((fake_t*)h, "init_params", 0, 0);
    return failing("init_params") ? -1 : 0;
}

static int record(void* h, int kind, int n, double sum, unsigned char* buf, int size, const char* name)
{
    fake_t* f = (fake_t*)h;
    if (f->log) { fprintf(f->log, "%s n=%d buf=%d\n", name, n, size); fflush(f->log); }
    if (failing("encode")) return -3;
    if (size < 16) return -1;
    float s = (float)sum;
    memcpy(buf, "FLAM", 4); memcpy(buf + 4, &kind, 4); memcpy(buf + 8, &n, 4); memcpy(buf + 12, &s, 4);
    return 16;
}
/* checksums: sum of |sample| over what the real entry point would read (both channels; left only when mono) */
int lame_encode_buffer_interleaved(void* h, short* pcm, int n, unsigned char* buf, int size)
{
    double s = 0; for (int i = 0; i < 2 * n; i++) s += fabs((double)pcm[i]);
    return record(h, 1, n, s, buf, size, "encode_buffer_interleaved");
}
int lame_encode_buffer(void* h, const short* l, const short* r, int n, unsigned char* buf, int size)
{
    double s = 0; for (int i = 0; i < n; i++) s += fabs((double)l[i]) + (((fake_t*)h)->channels == 2 ? fabs((double)r[i]) : 0.0);
    return record(h, 2, n, s, buf, size, "encode_buffer");
}
int lame_encode_buffer_interleaved_int(void* h, const int* pcm, int n, unsigned char* buf, int size)
{
    double s = 0; for (int i = 0; i < 2 * n; i++) s += fabs((double)pcm[i]) / 65536.0;
    return record(h, 3, n, s, buf, size, "encode_buffer_interleaved_int");
}
int lame_encode_buffer_int(void* h, const int* l, const int* r, int n, unsigned char* buf, int size)
{
    double s = 0; for (int i = 0; i < n; i++) s += (fabs((double)l[i]) + (((fake_t*)h)->channels == 2 ? fabs((double)r[i]) : 0.0)) / 65536.0;
    return record(h, 4, n, s, buf, size, "encode_buffer_int");
}
int lame_encode_buffer_interleaved_ieee_float(void* h, const float* pcm, int n, unsigned char* buf, int size)
{
    double s = 0; for (int i = 0; i < 2 * n; i++) s += fabs((double)pcm[i]);
    return record(h, 5, n, s, buf, size, "encode_buffer_interleaved_ieee_float");
}
int lame_encode_buffer_ieee_float(void* h, const float* l, const float* r, int n, unsigned char* buf, int size)
{
    double s = 0; for (int i = 0; i < n; i++) s += fabs((double)l[i]) + (((fake_t*)h)->channels == 2 ? fabs((double)r[i]) : 0.0);
    return record(h, 6, n, s, buf, size, "encode_buffer_ieee_float");
}

/* fake_soundtouch.c -- a STAND-IN for libSoundTouchDll with SoundTouchDLL's C entry points, implemented on the
 * oracle's own streaming model (oracle/nodey_oracle.c, orc_st_*).  It exists so that the pin harness
 * (oracle/real_soundtouch.py, tests/golden/make_st_golden.py, tests/test_st_real.py) can be exercised in an image
 * that has no SoundTouch: binding, driving loops, fixture round trip.  Comparing the oracle with this library proves
 * NOTHING about parity with SoundTouch -- it is the oracle on both sides -- and every test that uses it says so. */
#include <stdint.h>
#include <stdlib.h>

#include "../../oracle/nodey_oracle.h"

typedef struct { orc_st* st; int rate_hz, ch; float rate, pitch; } fake;

static void ensure(fake* f)
{
    if (!f->st) f->st = orc_st_create(f->rate_hz, f->ch, f->rate, f->pitch);
}

void* soundtouch_createInstance(void)
{
    fake* f = (fake*)calloc(1, sizeof(fake));
    f->rate_hz = 44100; f->ch = 2; f->rate = 1.0f; f->pitch = 1.0f;
    return f;
}
void soundtouch_destroyInstance(void* h) { fake* f = (fake*)h; if (!f) return; orc_st_destroy(f->st); free(f); }
const char* soundtouch_getVersionString(void) { return "fake (oracle model of 2.3.2)"; }
unsigned int soundtouch_getVersionId(void) { return 0; }      /* a real 2.3.2 returns 20302 */
void soundtouch_setRate(void* h, float v) { ((fake*)h)->rate = v; }
void soundtouch_setTempo(void* h, float v) { (void)h; (void)v; }
void soundtouch_setPitch(void* h, float v) { ((fake*)h)->pitch = v; }
int soundtouch_setChannels(void* h, unsigned int n) { ((fake*)h)->ch = (int)n; return 1; }
int soundtouch_setSampleRate(void* h, unsigned int r) { ((fake*)h)->rate_hz = (int)r; return 1; }
void soundtouch_flush(void* h) { fake* f = (fake*)h; ensure(f); orc_st_flush(f->st); }
int soundtouch_putSamples(void* h, const float* x, unsigned int n) { fake* f = (fake*)h; ensure(f); orc_st_put(f->st, x, n); return 1; }
unsigned int soundtouch_receiveSamples(void* h, float* out, unsigned int max) { fake* f = (fake*)h; ensure(f); return (unsigned int)orc_st_receive(f->st, out, max); }
unsigned int soundtouch_numSamples(void* h) { fake* f = (fake*)h; ensure(f); return (unsigned int)orc_st_num_samples(f->st); }
int soundtouch_getSetting(void* h, int id) { (void)h; (void)id; return -1; }

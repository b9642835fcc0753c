// st_shim.cpp -- SoundTouchDLL's C entry points over a plain libSoundTouch (C++ API), for oracle/real_soundtouch.py.
// TEST INFRASTRUCTURE ONLY.  Not built by default: the SoundTouch headers and library are not in this image.
//   make -C oracle st_shim SOUNDTOUCH_INC=/usr/include/soundtouch SOUNDTOUCH_LIB=-lSoundTouch
//   NODEY_REAL_SOUNDTOUCH=$PWD/oracle/_ref/libnodey_st_shim.so python -m pytest tests/test_st_real.py
// Same names and argument meaning as source/SoundTouchDLL/SoundTouchDLL.h of the SoundTouch distribution.
#include <SoundTouch.h>

using soundtouch::SoundTouch;

extern "C" {
void* soundtouch_createInstance() { return new SoundTouch(); }
void soundtouch_destroyInstance(void* h) { delete static_cast<SoundTouch*>(h); }
const char* soundtouch_getVersionString() { return SoundTouch::getVersionString(); }
unsigned int soundtouch_getVersionId() { return SoundTouch::getVersionId(); }
void soundtouch_setRate(void* h, float v) { static_cast<SoundTouch*>(h)->setRate(v); }
void soundtouch_setTempo(void* h, float v) { static_cast<SoundTouch*>(h)->setTempo(v); }
void soundtouch_setPitch(void* h, float v) { static_cast<SoundTouch*>(h)->setPitch(v); }
int soundtouch_setChannels(void* h, unsigned int n) { static_cast<SoundTouch*>(h)->setChannels(n); return 1; }
int soundtouch_setSampleRate(void* h, unsigned int r) { static_cast<SoundTouch*>(h)->setSampleRate(r); return 1; }
void soundtouch_flush(void* h) { static_cast<SoundTouch*>(h)->flush(); }
int soundtouch_putSamples(void* h, const float* x, unsigned int n) { static_cast<SoundTouch*>(h)->putSamples(x, n); return 1; }
unsigned int soundtouch_receiveSamples(void* h, float* out, unsigned int max) { return static_cast<SoundTouch*>(h)->receiveSamples(out, max); }
unsigned int soundtouch_numSamples(void* h) { return static_cast<SoundTouch*>(h)->numSamples(); }
int soundtouch_getSetting(void* h, int id) { return static_cast<SoundTouch*>(h)->getSetting(id); }
}

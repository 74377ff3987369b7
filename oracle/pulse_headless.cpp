// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
//
// Headless front end around the reference's UI controller: the UNMODIFIED Pulse.cpp (reference src/common/Pulse.cpp) driven
// through a BasePlatformInterface stub (reference BasePlatformInterface.h:13-19) instead of X11 / Win32 — what
// src/linux/main.cpp:234-287 does, without a display.  The same source is built twice:
//   * oracle/_ref/ref_pulse_headless   Pulse.cpp + the reference's own Render/Scene/... (oracle/Makefile, sources where they lie)
//   * build/shim_pulse_headless        Pulse.cpp (unchanged) against reflaxman_b200/shim + libreflax_b200.so: the GPU path
// and tests/test_shim.py compares what the two produce (front-end swap-in).
//
// Everything the controller sees is scripted and deterministic: a virtual performance counter that advances by a fixed
// pseudo-random pattern of 2 / 8 / 25 ms per query (so Pulse's adaptive chunk size both grows and shrinks,
// Pulse.cpp:136-145,183-194, identically in both builds), a fixed system time (the screenshot's file name, Pulse.cpp:156-172), a
// fixed window size, scripted key events, and rand() pinned (trace_math.h:34 seeds the LCGs from it).
//
//   pulse_headless OUTDIR/ [W H TICKS RES_KEY SS_KEY]
//     phase 1: onResize(W, H), TICKS calls of Pulse::exec with keys W / LEFT / SPACE pressed and released on a schedule; after
//              every completed frame the window is "repainted" exactly as linux/main.cpp:82-86 does (getRenderImagePixel for
//              every pixel, imageReady = false).  The last repaint is written to OUTDIR/interactive.bin (W*H uint32, row 0 = bottom)
//              and the HUD text to OUTDIR/hud.txt.
//     phase 2: F2, resolution key, sampling key (Pulse.cpp:383-441), exec until the controller is back in camera control:
//              Pulse::scrnshotRenderBegin / screenshotRenderProceed / screenshotRenderSave write OUTDIR/scrnshoot_<time>.bmp.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <stdint.h>

#include "Pulse.h"

extern "C" int rand(void) noexcept
{
  const char * s = getenv("RFX_SEED");
  return s ? (int)strtol(s, NULL, 0) : 12345;
}

namespace
{
  class HeadlessPlatform : public BasePlatformInterface
  {
  public:
    std::string dir;
    unsigned width, height;
    uint64_t counter;
    uint32_t lcg;
    unsigned invalidations, sleeps;

    HeadlessPlatform(const std::string & d, unsigned w, unsigned h) : dir(d), width(w), height(h), counter(1000), lcg(2024), invalidations(0), sleeps(0) {}
    std::string getExePath() { return dir; }
    uint64_t getPerformanceCounter()
    {
      lcg = lcg * 1664525u + 1013904223u;
      const unsigned k = (lcg >> 24) % 8u;
      counter += k == 0 ? 25 : (k < 3 ? 8 : 2);     // milliseconds of virtual time per query
      return counter;
    }
    uint64_t getPerformanceFrequency() { return 1000; }
    uint64_t getSystemTime() { return 0x0000000100000002ull; }
    void getMainWindowClientSize(unsigned int * const w, unsigned int * const h) { *w = width; *h = height; }
    void invalidateMainWindow() { invalidations++; }
    void sleep(unsigned int) { sleeps++; }
  };

  KEY_CODE digitKey(int d) { static const KEY_CODE k[9] = { KEY_1, KEY_2, KEY_3, KEY_4, KEY_5, KEY_6, KEY_7, KEY_8, KEY_9 }; return k[d - 1]; }
}

int main(int argc, char ** argv)
{
  if (argc < 2) { fprintf(stderr, "usage: pulse_headless OUTDIR/ [W H TICKS RES_KEY SS_KEY]\n"); return 2; }
  const std::string dir = argv[1];
  const unsigned W = argc > 2 ? atoi(argv[2]) : 160, H = argc > 3 ? atoi(argv[3]) : 120;
  const int ticks = argc > 4 ? atoi(argv[4]) : 260;
  const int resKey = argc > 5 ? atoi(argv[5]) : 2, ssKey = argc > 6 ? atoi(argv[6]) : 2;   // 1024x768, 2x2 SSAA

  HeadlessPlatform plat(dir, W, H);
  Pulse pulse(&plat);
  pulse.exec();                       // stInit: sleeps (Pulse.cpp:231-233)
  pulse.onResize(W, H);               // first ConfigureNotify: stInit -> stCameraControl + setImageSize (Pulse.cpp:443-453)

  // ---- phase 1: interactive loop with scripted keys
  struct Ev { int tick; KEY_CODE key; bool down; };
  const Ev script[] = { { 3, KEY_W, true }, { 14, KEY_W, false }, { 16, KEY_LEFT, true }, { 30, KEY_LEFT, false }, { 34, KEY_SPACE, true },
                        { 40, KEY_SPACE, false }, { 41, KEY_D, true }, { 47, KEY_A, true }, { 52, KEY_D, false }, { 60, KEY_A, false } };
  std::vector<uint32_t> window(size_t(W) * H, 0);
  int repaints = 0;
  for (int t = 0; t < ticks; t++)
  {
    for (size_t k = 0; k < sizeof(script) / sizeof(script[0]); k++)
      if (script[k].tick == t) pulse.onKeyEvent(script[k].key, script[k].down);
    pulse.exec();
    if (pulse.imageReady)             // Expose: linux/main.cpp:55-87
    {
      for (unsigned y = 0; y < H; y++)
        for (unsigned x = 0; x < W; x++)
          window[x + size_t(y) * W] = pulse.getRenderImagePixel(x, y);
      pulse.imageReady = false;
      repaints++;
    }
  }
  {
    FILE * f = fopen((dir + "interactive.bin").c_str(), "wb");
    if (!f) return 3;
    fwrite(window.data(), 4, window.size(), f);
    fclose(f);
    f = fopen((dir + "hud.txt").c_str(), "w");
    if (!f) return 3;
    std::vector<std::string> * text = pulse.getCurrentScreenText();
    for (size_t i = 0; i < text->size(); i++)
      if ((*text)[i].compare(0, 10, "Frame time") != 0) fprintf(f, "%s\n", (*text)[i].c_str());
    fprintf(f, "repaints %d\n", repaints);
    fclose(f);
  }

  // ---- phase 2: the screenshot flow
  pulse.onKeyEvent(KEY_F2, true); pulse.onKeyEvent(KEY_F2, false);
  pulse.exec();                                                        // resolution menu: sleeps
  pulse.onKeyEvent(digitKey(resKey), true); pulse.onKeyEvent(digitKey(resKey), false);
  pulse.onKeyEvent(digitKey(ssKey), true); pulse.onKeyEvent(digitKey(ssKey), false);
  int guard = 0;
  for (;; guard++)
  {
    pulse.exec();
    const std::vector<std::string> * text = pulse.getCurrentScreenText();
    if (!text->empty() && (*text)[0].compare(0, 12, "Resolution :") == 0) break;     // back in stCameraControl
    if (guard > 100000000) return 4;
  }
  printf("{\"repaints\": %d, \"screenshot_exec_calls\": %d, \"invalidations\": %u, \"sleeps\": %u, \"virtual_ms\": %llu}\n",
         repaints, guard + 1, plat.invalidations, plat.sleeps, (unsigned long long)plat.counter);
  return 0;
}

/* Headless stand-in for <SDL.h>, used ONLY to compile the unmodified reference
 * renderers (raytracer/Source/skeleton.cpp, rasteriser/Source/skeleton.cpp and
 * their SDLauxiliary.h) into oracle/_ref/ as a test oracle.  SDL2 is not
 * installed in the build image and windowing is not on the hot path.
 *
 * Every entry point is an inert no-op: nothing is drawn, no events arrive.
 * Test infrastructure -- never linked into the product library. */
#ifndef B200_ORACLE_SDL_STUB_H
#define B200_ORACLE_SDL_STUB_H

#include <stdint.h>
#include <string.h>
#include <stdlib.h>

typedef struct SDL_Window { int unused; } SDL_Window;
typedef struct SDL_Renderer { int unused; } SDL_Renderer;
typedef struct SDL_Texture { int unused; } SDL_Texture;
typedef struct SDL_Surface { void *pixels; int w, h, pitch; } SDL_Surface;
typedef struct SDL_version { uint8_t major, minor, patch; } SDL_version;

typedef struct SDL_Keysym { int sym; } SDL_Keysym;
typedef struct SDL_KeyboardEvent { SDL_Keysym keysym; } SDL_KeyboardEvent;
typedef struct SDL_Event { uint32_t type; SDL_KeyboardEvent key; } SDL_Event;

#define SDL_VERSION(v) do { (v)->major = 2; (v)->minor = 0; (v)->patch = 0; } while (0)
#define SDL_BIG_ENDIAN 4321
#define SDL_LIL_ENDIAN 1234
#define SDL_BYTEORDER SDL_LIL_ENDIAN

#define SDL_INIT_TIMER 0x1u
#define SDL_INIT_VIDEO 0x20u
#define SDL_WINDOW_OPENGL 0x2u
#define SDL_WINDOW_FULLSCREEN_DESKTOP 0x1001u
#define SDL_WINDOWPOS_UNDEFINED 0x1FFF0000
#define SDL_RENDERER_ACCELERATED 0x2u
#define SDL_RENDERER_PRESENTVSYNC 0x4u
#define SDL_HINT_RENDER_SCALE_QUALITY "SDL_RENDER_SCALE_QUALITY"
#define SDL_PIXELFORMAT_ARGB8888 0x16362004u
#define SDL_TEXTUREACCESS_STATIC 0

#define SDL_QUIT 0x100u
#define SDL_KEYDOWN 0x300u

enum {
  SDLK_ESCAPE = 27, SDLK_SPACE = ' ', SDLK_1 = '1', SDLK_2 = '2',
  SDLK_a = 'a', SDLK_d = 'd', SDLK_e = 'e', SDLK_f = 'f', SDLK_g = 'g',
  SDLK_i = 'i', SDLK_m = 'm', SDLK_n = 'n', SDLK_o = 'o', SDLK_q = 'q',
  SDLK_s = 's', SDLK_w = 'w', SDLK_x = 'x', SDLK_z = 'z',
  SDLK_RIGHT = 0x4000004F, SDLK_LEFT = 0x40000050,
  SDLK_DOWN = 0x40000051, SDLK_UP = 0x40000052
};

static inline int SDL_Init(uint32_t) { return 0; }
static inline void SDL_Quit(void) {}
static inline const char *SDL_GetError(void) { return "headless SDL stub"; }
static inline void SDL_GetVersion(SDL_version *v) { SDL_VERSION(v); }
static inline SDL_Window *SDL_CreateWindow(const char *, int, int, int, int, uint32_t) {
  static SDL_Window w; return &w;
}
static inline SDL_Renderer *SDL_CreateRenderer(SDL_Window *, int, uint32_t) {
  static SDL_Renderer r; return &r;
}
static inline SDL_Texture *SDL_CreateTexture(SDL_Renderer *, uint32_t, int, int, int) {
  static SDL_Texture t; return &t;
}
static inline int SDL_SetHint(const char *, const char *) { return 1; }
static inline int SDL_RenderSetLogicalSize(SDL_Renderer *, int, int) { return 0; }
static inline void SDL_DestroyTexture(SDL_Texture *) {}
static inline void SDL_DestroyRenderer(SDL_Renderer *) {}
static inline void SDL_DestroyWindow(SDL_Window *) {}
static inline int SDL_UpdateTexture(SDL_Texture *, const void *, const void *, int) { return 0; }
static inline int SDL_RenderClear(SDL_Renderer *) { return 0; }
static inline int SDL_RenderCopy(SDL_Renderer *, SDL_Texture *, const void *, const void *) { return 0; }
static inline void SDL_RenderPresent(SDL_Renderer *) {}
static inline SDL_Surface *SDL_CreateRGBSurfaceFrom(void *px, int w, int h, int, int pitch,
                                                    uint32_t, uint32_t, uint32_t, uint32_t) {
  static SDL_Surface s; s.pixels = px; s.w = w; s.h = h; s.pitch = pitch; return &s;
}
static inline uint32_t SDL_GetTicks(void) { return 0; }
#ifdef B200_SDL_SCRIPTED
/* Scripted mode (oracle/refbuild/prog_harness.cpp): the reference's own main() / Update() loop
 * runs headless.  The script is a list of frames, each a list of key codes: SDL_PollEvent hands
 * out the current frame's keys as SDL_KEYDOWN events, then returns 0 (Update() returns true and
 * main() draws the frame); once every frame has been drawn it delivers SDL_QUIT.  SDL_SaveBMP
 * (the screenshot main() takes after its loop, SDLauxiliary.h:24-53) copies the surface out. */
static const int *sdl_script_keys = 0;      /* all key codes, frame after frame */
static const int *sdl_script_lens = 0;      /* keys per frame */
static int sdl_script_frames = 0, sdl_script_frame = 0, sdl_script_key = 0, sdl_script_at = 0;
static uint32_t *sdl_script_shot = 0;       /* receives the screenshot */
static inline int SDL_PollEvent(SDL_Event *e) {
  if (sdl_script_frame >= sdl_script_frames) { e->type = SDL_QUIT; return 1; }
  if (sdl_script_key < sdl_script_lens[sdl_script_frame]) {
    e->type = SDL_KEYDOWN;
    e->key.keysym.sym = sdl_script_keys[sdl_script_at++];
    ++sdl_script_key;
    return 1;
  }
  ++sdl_script_frame; sdl_script_key = 0;
  return 0;
}
static inline int SDL_SaveBMP(SDL_Surface *s, const char *) {
  if (sdl_script_shot) memcpy(sdl_script_shot, s->pixels, (size_t)s->pitch * (size_t)s->h);
  return 0;
}
#else
static inline int SDL_SaveBMP(SDL_Surface *, const char *) { return 0; }
static inline int SDL_PollEvent(SDL_Event *) { return 0; }
#endif

#endif

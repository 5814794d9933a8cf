// TEST INFRASTRUCTURE.  The reference's own Sigma-Delta implementation (package_bgs/bl/sdLaMa091.cpp, plain C in a .cpp
// file, no OpenCV), compiled from where it lies under /root/reference into oracle/_ref/libdp_ref.so and driven as
// SigmaDeltaBGS::process does (package_bgs/bl/SigmaDeltaBGS.cpp:21-50): parameters applied before every frame, the first
// frame only initialises (no output), later frames update and the first channel of the 3-channel map is the mask.
//
// One thing is pinned down here that the reference leaves undefined: sdLaMa091AllocInit_8u_C3R delegates to the C1R
// initialiser, which fills only the first `width` BYTES of every row of the variance image Vt with Vmin and leaves the
// other two thirds as malloc() returned them (sdLaMa091.cpp:186-196, :204-216).  For frames of camera size those blocks
// come fresh from mmap, i.e. zero-filled; this build makes that the rule by giving the file calloc() for malloc().
#include <stdlib.h>
#include <string.h>
static inline void *sd_ref_zero_alloc(size_t n) { return calloc(1, n); }
#define malloc(n) sd_ref_zero_alloc(n)
#include "sdLaMa091.cpp"
#undef malloc

struct sd_ref { sdLaMa091_t *alg; int w, h; bool first; unsigned char *tmp; };

extern "C" {
#define SD_EXPORT __attribute__((visibility("default")))
SD_EXPORT sd_ref *sd_ref_create(int w, int h)
{
    sd_ref *r = new sd_ref;
    r->alg = sdLaMa091New(); r->w = w; r->h = h; r->first = true;
    r->tmp = (unsigned char *)calloc(1, (size_t)w * h * 3);
    return r;
}
// returns 1 and fills fg (h * w bytes) when the plugin produced a mask, 0 on the first frame
SD_EXPORT int sd_ref_process(sd_ref *r, const unsigned char *bgr, unsigned char *fg, int ampFactor, int minVar, int maxVar)
{
    sdLaMa091SetAmplificationFactor(r->alg, ampFactor);               // loadConfig -> applyParams, every frame (:26, :74, :78-82)
    sdLaMa091SetMinimalVariance(r->alg, minVar);
    sdLaMa091SetMaximalVariance(r->alg, maxVar);
    if (r->first) {                                                    // :28-33
        sdLaMa091AllocInit_8u_C3R(r->alg, bgr, r->w, r->h, r->w * 3);
        r->first = false;
        return 0;
    }
    sdLaMa091Update_8u_C3R(r->alg, bgr, r->tmp);                       // :37
    for (size_t i = 0; i < (size_t)r->w * r->h; i++) fg[i] = r->tmp[3 * i];      // :39-45
    return 1;
}
SD_EXPORT void sd_ref_destroy(sd_ref *r) { sdLaMa091Free(r->alg); free(r->tmp); delete r; }
}

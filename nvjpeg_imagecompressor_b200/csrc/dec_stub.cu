// temporary decoder stub
#include "dec.h"
namespace b2j {
struct Decoder { int x; };
Decoder *dec_create(int, char *, size_t) { return new Decoder(); }
void dec_destroy(Decoder *d) { delete d; }
int dec_run(Decoder *, const uint8_t *, size_t, const JpegInfo &, const Geom &, uint8_t *, size_t, cudaStream_t, b2j_timings *, uint64_t *) { return B2J_EINTERNAL; }
int dec_check(Decoder *, char *, size_t) { return B2J_OK; }
const void *dec_coef_ptr(Decoder *, size_t *n) { *n = 0; return nullptr; }
}

// ref_hdrloader_shim.cpp -- C entry points around the REFERENCE's own HDRLoader::load
// (P5/lib/hdrloader.cpp, compiled where it lies under /root/reference into oracle/_ref/;
// see ezrt_b200/build.py:build_reference_hdrloader).  Test infrastructure only: it lets
// tests/test_hdr.py compare ezrt_hdr_load() with the unmodified reference decoder.
#include <string.h>

#include "hdrloader.h"

extern "C" {

// returns 0 on success; *cols is owned by the reference (new[]), release with ref_hdr_free
int ref_hdr_load(const char* path, int* width, int* height, float** cols) {
    HDRLoaderResult res;
    res.width = res.height = 0;
    res.cols = 0;
    if (!HDRLoader::load(path, res)) return -1;
    *width = res.width;
    *height = res.height;
    *cols = res.cols;
    return 0;
}

void ref_hdr_copy(const float* cols, float* dst, size_t n_floats) { memcpy(dst, cols, n_floats * sizeof(float)); }
void ref_hdr_free(float* cols) { delete[] cols; }

}

// The reference parses the resolution line with sscanf("-Y %ld +X %ld", &h, &w) into two
// `int`s (P5/lib/hdrloader.cpp:67-68): on LP64 each %ld stores 8 bytes, so the second store
// zeroes `h`.  The build recipe therefore compiles the (unmodified) source with
// -Dsscanf=ezrt_ref_sscanf and this function supplies what the call means on its original
// ILP32/LLP64 target: two decimal ints.
#include <stdarg.h>
#include <stdlib.h>

extern "C" int ezrt_ref_sscanf(const char* s, const char* fmt, ...) {
    (void)fmt;
    va_list ap;
    va_start(ap, fmt);
    int* h = va_arg(ap, int*);
    int* w = va_arg(ap, int*);
    va_end(ap);
    if (s[0] != '-' || s[1] != 'Y') return 0;
    char* end = 0;
    long a = strtol(s + 2, &end, 10);
    if (end == s + 2) return 0;
    *h = (int)a;
    while (*end == ' ') end++;
    if (end[0] != '+' || end[1] != 'X') return 1;
    char* end2 = 0;
    long b = strtol(end + 2, &end2, 10);
    if (end2 == end + 2) return 1;
    *w = (int)b;
    return 2;
}

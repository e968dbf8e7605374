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

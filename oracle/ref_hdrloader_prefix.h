// Force-included (-include) in front of the reference's hdrloader.cpp by the oracle/_ref build
// recipe: routes its one sscanf("-Y %ld +X %ld", int*, int*) call (UB on LP64, zeroes the height)
// to ezrt_ref_sscanf() in ref_hdrloader_shim.cpp.  The reference source itself is not modified.
#include <stdio.h>
#ifdef __cplusplus
extern "C"
#endif
int ezrt_ref_sscanf(const char* s, const char* fmt, ...);
#define sscanf ezrt_ref_sscanf

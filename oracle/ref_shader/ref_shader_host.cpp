// ref_shader_host.cpp -- runs the reference's own fragment shaders (transpiled source text) on the CPU.
//
// *** TEST INFRASTRUCTURE (oracle/), NOT PRODUCT.  Builds into oracle/_ref/libezrt_refshader.so, only
// *** where /root/reference exists; tests use it to pin oracle/ezrt_oracle.cpp and to generate the
// *** golden frames under tests/golden/ (tests/golden/make_golden_refshader.py).
//
// The four `shader_<variant>.inc` files are the text of  P3|P4|P5/shaders/fshader.fsh  after the
// mechanical edits listed in transpile.py; nothing of them is committed.  This file is the "GL side"
// a fragment needs: the uniforms (P5/main.cpp:876-905 sets them per frame), the full-screen quad's
// interpolated `pix` (vshader.vsh: pix = vPosition in [-1,1]^2), and the lastFrame ping-pong
// (pass1 -> pass2 copies the frame into lastFrame, P5/main.cpp:907-911).
#include <stdint.h>
#include <string.h>

#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

#include "ezrt.h"
#include "glsl_emul.h"

namespace glsl {

#define EZRT_SHADER_STRUCT(NAME, INC)                                            \
    struct NAME {                                                                \
        const Uniforms& U;                                                       \
        vec3 pix_in;                                                             \
        int max_bounce_in;                                                       \
        vec4 gl_FragData[1];                                                     \
        NAME(const Uniforms& u, vec3 p, int mb) : U(u), pix_in(p), max_bounce_in(mb) {}

// clang-format off
EZRT_SHADER_STRUCT(Shader_p3, 0)
#include "shader_p3.inc"
};
#undef PI
#undef INF
#undef SIZE_TRIANGLE
#undef SIZE_BVHNODE

EZRT_SHADER_STRUCT(Shader_p4, 0)
#include "shader_p4.inc"
};
#undef PI
#undef INF
#undef SIZE_TRIANGLE
#undef SIZE_BVHNODE

EZRT_SHADER_STRUCT(Shader_p5sobol, 0)
#include "shader_p5sobol.inc"
};
#undef PI
#undef INF
#undef SIZE_TRIANGLE
#undef SIZE_BVHNODE

EZRT_SHADER_STRUCT(Shader_p5is, 0)
#include "shader_p5is.inc"
};
#undef PI
#undef INF
#undef SIZE_TRIANGLE
#undef SIZE_BVHNODE
// clang-format on

EZRT_SHADER_STRUCT(Shader_pass3, 0)
#include "shader_pass3.inc"
};

template <class S>
static vec3 run_fragment(const Uniforms& U, vec3 pix, int max_bounce) {
    S s(U, pix, max_bounce);
    s.shader_main();
    return s.gl_FragData[0].rgb();
}

}  // namespace glsl

extern "C" {

// Same arguments and framebuffer convention as oracle_render (row 0 = bottom row, py = 0):
// renders frames first_frame .. first_frame+spp-1 of the shader that implements p->mode, every frame
// a full pass over the image that reads the previous frame through the lastFrame sampler.
int refshader_render(const float* tris, int nTriangles, const float* nodes, int nNodes, const float* hdr,
                     const float* hdrCache, int hdrW, int hdrH, int hdrLinear, const ezrt_render_params* p,
                     float* framebuffer, int n_threads) {
    using namespace glsl;
    if (!tris || !nodes || !p || !framebuffer || !hdr) return -1;
    if (p->mode == EZRT_MODE_DISNEY_IS_MIS_P5 && !hdrCache) return -1;
    const int W = p->width, H = p->height, C = (p->out_channels == 4) ? 4 : 3;
    std::vector<float> last((size_t)W * H * 3), next((size_t)W * H * 3);
    for (size_t i = 0; i < (size_t)W * H; i++)
        for (int k = 0; k < 3; k++) last[3 * i + k] = (p->first_frame == 0) ? 0.0f : framebuffer[i * C + k];
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
    for (int s = 0; s < p->spp; s++) {
        Uniforms U;
        U.frameCounter = p->first_frame + (uint32_t)s;
        U.nTriangles = nTriangles;
        U.nNodes = nNodes;
        U.width = W;
        U.height = H;
        U.hdrResolution = hdrW;  // P5/main.cpp:842 (hdrResolution = hdrRes.width)
        U.triangles.texels = tris;
        U.nodes.texels = nodes;
        U.lastFrame = sampler2D{last.data(), W, H, 3, 0};
        U.hdrMap = sampler2D{hdr, hdrW, hdrH, 3, hdrLinear ? 1 : 0};
        U.hdrCache = sampler2D{hdrCache, hdrW, hdrH, 3, hdrLinear ? 1 : 0};
        U.eye = vec3(p->eye[0], p->eye[1], p->eye[2]);
        memcpy(U.cameraRotate.m, p->camera_rotate, sizeof(float) * 16);
#pragma omp parallel for schedule(dynamic, 1)
        for (int py = 0; py < H; py++) {
            for (int px = 0; px < W; px++) {
                // the rasteriser's interpolation of vPosition at the pixel centre (oracle header, "pix")
                vec3 pix(((float)px + 0.5f) / (float)W * 2.0f - 1.0f, ((float)py + 0.5f) / (float)H * 2.0f - 1.0f, 0.0f);
                vec3 c;
                switch (p->mode) {
                    case EZRT_MODE_DIFFUSE_P3: c = run_fragment<Shader_p3>(U, pix, p->max_bounce); break;
                    case EZRT_MODE_DISNEY_ANISO_P4: c = run_fragment<Shader_p4>(U, pix, p->max_bounce); break;
                    case EZRT_MODE_DISNEY_SOBOL_P5: c = run_fragment<Shader_p5sobol>(U, pix, p->max_bounce); break;
                    default: c = run_fragment<Shader_p5is>(U, pix, p->max_bounce); break;
                }
                float* d = &next[((size_t)py * W + px) * 3];
                d[0] = c.x; d[1] = c.y; d[2] = c.z;
            }
        }
        last.swap(next);
    }
    for (size_t i = 0; i < (size_t)W * H; i++) {
        for (int k = 0; k < 3; k++) framebuffer[i * C + k] = last[3 * i + k];
        if (C == 4) framebuffer[i * C + 3] = 1.0f;
    }
    return 0;
}

// pass3.fsh (tone mapping + gamma) over a W x H image with `channels` floats per pixel; out: 3 floats per pixel.
// texPass0 is sampled at the pixel's own texel (GL_NEAREST stand-in for the screen-sized attachment).
int refshader_pass3(const float* in, int channels, int W, int H, float* out) {
    using namespace glsl;
    if (!in || !out || channels < 3) return -1;
    Uniforms U;
    memset(&U, 0, sizeof(U));
    U.texPass0 = sampler2D{in, W, H, channels, 0};
#pragma omp parallel for
    for (int py = 0; py < H; py++)
        for (int px = 0; px < W; px++) {
            vec3 pix(((float)px + 0.5f) / (float)W * 2.0f - 1.0f, ((float)py + 0.5f) / (float)H * 2.0f - 1.0f, 0.0f);
            Shader_pass3 s(U, pix, 0);
            s.shader_main();
            float* d = out + ((size_t)py * W + px) * 3;
            d[0] = s.fragColor.x; d[1] = s.fragColor.y; d[2] = s.fragColor.z;
        }
    return 0;
}

}  // extern "C"

// ref_host_shim.cpp -- calls the reference's own host code (main.cpp of part 3, 4 or 5 -- EZRT_REF_PART --
// compiled from where it lies; the functions below are the same in all three parts, line numbers are P5's).
//
// *** TEST INFRASTRUCTURE (oracle/), NOT PRODUCT.  Builds into oracle/_ref/libezrt_refhost.so, only where
// *** /root/reference exists.  Nothing of the reference is copied: this translation unit #includes
// *** main.cpp by absolute path (-DEZRT_REF_MAIN_CPP=...), with main renamed (-Dmain=ezrt_ref_main) and the
// *** absent third-party headers replaced by the stand-ins in oracle/ref_stubs/ (GL calls: no-ops that
// *** capture uploads; glm: the subset main.cpp uses).
//
// What this pins (tests/test_ref_host.py): the product's readObj / buildBVH / buildBVHwithSAH / encode /
// calculateHdrCache restatements (ezrt_b200/csrc/host_scene.cpp) against the code they
// restate -- P5/main.cpp:274-392 readObj, :395-447 buildBVH, :450-588 buildBVHwithSAH, :591-683
// calculateHdrCache, :805-838 the encode loops, and main() as a whole (its shipped scene: teapot +
// 13000-unit floor + chinese_garden_2k.hdr), byte for byte.
#include EZRT_REF_MAIN_CPP

#include <unistd.h>

namespace {
std::vector<Triangle> g_tris;
std::vector<BVHNode> g_nodes;

Material material_from18(const float* m) {
    Material r;
    r.emissive = vec3(m[0], m[1], m[2]);
    r.baseColor = vec3(m[3], m[4], m[5]);
    r.subsurface = m[6]; r.metallic = m[7]; r.specular = m[8]; r.specularTint = m[9];
    r.roughness = m[10]; r.anisotropic = m[11]; r.sheen = m[12]; r.sheenTint = m[13];
    r.clearcoat = m[14]; r.clearcoatGloss = m[15]; r.IOR = m[16]; r.transmission = m[17];
    return r;
}
}  // namespace

extern "C" {

void refhost_reset() {
    g_tris.clear();
    g_nodes.clear();
}

// getTransformMatrix (P5/main.cpp:255-271), column-major out[16]
void refhost_transform_matrix(const float rot[3], const float tr[3], const float sc[3], float out[16]) {
    mat4 m = getTransformMatrix(vec3(rot[0], rot[1], rot[2]), vec3(tr[0], tr[1], tr[2]), vec3(sc[0], sc[1], sc[2]));
    memcpy(out, value_ptr(m), sizeof(float) * 16);
}

// readObj (P5/main.cpp:274-392) appended to the current triangle list; returns the list's size
int refhost_read_obj(const char* path, const float material[18], const float trans[16], int smooth) {
    mat4 m;
    memcpy(&m.c[0].x, trans, sizeof(float) * 16);
    readObj(path, g_tris, material_from18(material), m, smooth != 0);
    return (int)g_tris.size();
}

// main()'s "build bvh" block (P5/main.cpp:789-800): dummy node 0, then buildBVH (sah = 0) or buildBVHwithSAH
int refhost_build_bvh(int leaf_n, int sah) {
    BVHNode testNode;
    testNode.left = 255; testNode.right = 128; testNode.n = 30; testNode.index = 0;
    testNode.AA = vec3(1, 1, 0); testNode.BB = vec3(0, 1, 0);
    g_nodes.assign(1, testNode);
    if (sah) buildBVHwithSAH(g_tris, g_nodes, 0, (int)g_tris.size() - 1, leaf_n);
    else buildBVH(g_tris, g_nodes, 0, (int)g_tris.size() - 1, leaf_n);
    return (int)g_nodes.size();
}

int refhost_counts(int* n_tris, int* n_nodes) {
    *n_tris = (int)g_tris.size();
    *n_nodes = (int)g_nodes.size();
    return 0;
}

// the field-by-field copies of main()'s encode loops (P5/main.cpp:805-838) -- restated here because they
// live inside main(); refhost_run_main() below captures the reference's own
void refhost_encode(float* tris36, float* nodes12) {
    for (size_t i = 0; i < g_tris.size(); i++) {
        const Triangle& t = g_tris[i];
        const Material& m = t.material;
        const vec3 v[12] = {t.p1, t.p2, t.p3, t.n1, t.n2, t.n3, m.emissive, m.baseColor,
                            vec3(m.subsurface, m.metallic, m.specular), vec3(m.specularTint, m.roughness, m.anisotropic),
                            vec3(m.sheen, m.sheenTint, m.clearcoat), vec3(m.clearcoatGloss, m.IOR, m.transmission)};
        memcpy(tris36 + i * 36, v, sizeof(v));
    }
    for (size_t i = 0; i < g_nodes.size(); i++) {
        const vec3 v[4] = {vec3(g_nodes[i].left, g_nodes[i].right, 0), vec3(g_nodes[i].n, g_nodes[i].index, 0), g_nodes[i].AA, g_nodes[i].BB};
        memcpy(nodes12 + i * 12, v, sizeof(v));
    }
}

#if EZRT_REF_PART == 5
// calculateHdrCache (P5/main.cpp:591-683; parts 3 and 4 have no importance sampling)
void refhost_hdr_cache(const float* hdr, int w, int h, float* out) {
    float* c = calculateHdrCache(const_cast<float*>(hdr), w, h);
    memcpy(out, c, sizeof(float) * 3 * (size_t)w * h);
    delete[] c;
}
#endif

// Runs the reference's main() in `source_dir` (it opens models/, HDR/ and shaders/ relative to the cwd) up to
// glutMainLoop(), which the stand-in returns from.  Returns the number of captured uploads; in call order they
// are: triangle texture buffer, BVH texture buffer, HDR map and (part 5) HDR sampling cache (P5/main.cpp:843-868).
int refhost_run_main(const char* source_dir) {
    char cwd[4096];
    if (!getcwd(cwd, sizeof(cwd))) return -1;
    if (chdir(source_dir) != 0) return -2;
    ezrt_gl_capture().uploads.clear();
    char arg0[] = "ezrt_ref_main";
    char* argv[] = {arg0, nullptr};
    ezrt_ref_main(1, argv);
    if (chdir(cwd) != 0) return -3;
    return (int)ezrt_gl_capture().uploads.size();
}
long long refhost_upload_info(int i, int* width, int* height, unsigned* target) {
    const ezrt_gl_upload& u = ezrt_gl_capture().uploads.at(i);
    *width = u.width; *height = u.height; *target = u.target;
    return (long long)u.bytes.size();
}
void refhost_upload_copy(int i, void* dst) {
    const ezrt_gl_upload& u = ezrt_gl_capture().uploads.at(i);
    memcpy(dst, u.bytes.data(), u.bytes.size());
}

}  // extern "C"

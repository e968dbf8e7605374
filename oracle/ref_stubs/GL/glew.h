// GL/glew.h -- STAND-IN for the OpenGL entry points the reference's main.cpp calls.
//
// *** TEST INFRASTRUCTURE (oracle/), NOT PRODUCT.  There is no GL here: every call is a no-op, except
// *** that the uploads main() makes -- glBufferData(GL_TEXTURE_BUFFER, ...) for the triangle and BVH
// *** texture buffers, glTexImage2D(..., data) for the HDR map and its sampling cache -- are CAPTURED
// *** (ezrt_gl_capture) so a test can see exactly what the reference would hand to the GPU.
#ifndef EZRT_STUB_GLEW_H
#define EZRT_STUB_GLEW_H

#include <stddef.h>
#include <string.h>

#include <vector>

typedef unsigned int GLuint;
typedef int GLint;
typedef int GLsizei;
typedef unsigned int GLenum;
typedef unsigned int GLbitfield;
typedef unsigned char GLboolean;
typedef float GLfloat;
typedef char GLchar;
typedef void GLvoid;
typedef ptrdiff_t GLsizeiptr;
typedef ptrdiff_t GLintptr;

enum {
    GL_FALSE = 0, GL_TRUE = 1, GL_TRIANGLES = 4, GL_DEPTH_BUFFER_BIT = 0x100, GL_COLOR_BUFFER_BIT = 0x4000,
    GL_DEPTH_TEST = 0x0B71, GL_TEXTURE_2D = 0x0DE1, GL_FLOAT = 0x1406, GL_RGB = 0x1907, GL_RGBA = 0x1908,
    GL_NEAREST = 0x2600, GL_LINEAR = 0x2601, GL_TEXTURE_MAG_FILTER = 0x2800, GL_TEXTURE_MIN_FILTER = 0x2801, GL_TEXTURE_WRAP_S = 0x2802,
    GL_TEXTURE_WRAP_T = 0x2803, GL_CLAMP_TO_EDGE = 0x812F, GL_TEXTURE0 = 0x84C0, GL_TEXTURE1, GL_TEXTURE2, GL_TEXTURE3,
    GL_TEXTURE4, GL_RGBA32F = 0x8814, GL_RGB32F = 0x8815, GL_ARRAY_BUFFER = 0x8892, GL_STATIC_DRAW = 0x88E4,
    GL_FRAGMENT_SHADER = 0x8B30, GL_VERTEX_SHADER = 0x8B31, GL_COMPILE_STATUS = 0x8B81, GL_TEXTURE_BUFFER = 0x8C2A,
    GL_COLOR_ATTACHMENT0 = 0x8CE0, GL_FRAMEBUFFER = 0x8D40
};

struct ezrt_gl_upload {
    GLenum target, internal_format;
    int width, height;          // 2D textures; 0 for buffers
    std::vector<unsigned char> bytes;
};
struct ezrt_gl_capture_t {
    std::vector<ezrt_gl_upload> uploads;
    GLuint next_name = 1;
};
inline ezrt_gl_capture_t& ezrt_gl_capture() {
    static ezrt_gl_capture_t c;
    return c;
}

inline void glGenTextures(GLsizei n, GLuint* p) { for (int i = 0; i < n; i++) p[i] = ezrt_gl_capture().next_name++; }
inline void glGenBuffers(GLsizei n, GLuint* p) { for (int i = 0; i < n; i++) p[i] = ezrt_gl_capture().next_name++; }
inline void glGenVertexArrays(GLsizei n, GLuint* p) { for (int i = 0; i < n; i++) p[i] = ezrt_gl_capture().next_name++; }
inline void glGenFramebuffers(GLsizei n, GLuint* p) { for (int i = 0; i < n; i++) p[i] = ezrt_gl_capture().next_name++; }
inline void glBindTexture(GLenum, GLuint) {}
inline void glBindBuffer(GLenum, GLuint) {}
inline void glBindVertexArray(GLuint) {}
inline void glBindFramebuffer(GLenum, GLuint) {}
inline void glActiveTexture(GLenum) {}
inline void glTexParameteri(GLenum, GLenum, GLint) {}
inline void glTexBuffer(GLenum, GLenum, GLuint) {}
inline void glBufferData(GLenum target, GLsizeiptr size, const void* data, GLenum) {
    if (target != GL_TEXTURE_BUFFER || !data) return;
    ezrt_gl_upload u;
    u.target = target; u.internal_format = 0; u.width = u.height = 0;
    u.bytes.assign((const unsigned char*)data, (const unsigned char*)data + size);
    ezrt_gl_capture().uploads.push_back(u);
}
inline void glBufferSubData(GLenum, GLintptr, GLsizeiptr, const void*) {}
inline void glTexImage2D(GLenum target, GLint, GLint internal_format, GLsizei w, GLsizei h, GLint, GLenum format, GLenum, const void* data) {
    if (!data) return;
    ezrt_gl_upload u;
    u.target = target; u.internal_format = (GLenum)internal_format; u.width = w; u.height = h;
    size_t size = (size_t)w * h * (format == GL_RGBA ? 4 : 3) * sizeof(float);
    u.bytes.assign((const unsigned char*)data, (const unsigned char*)data + size);
    ezrt_gl_capture().uploads.push_back(u);
}
inline void glEnableVertexAttribArray(GLuint) {}
inline void glVertexAttribPointer(GLuint, GLint, GLenum, GLboolean, GLsizei, const void*) {}
inline void glFramebufferTexture2D(GLenum, GLenum, GLenum, GLuint, GLint) {}
inline void glDrawBuffers(GLsizei, const GLuint*) {}
inline void glUseProgram(GLuint) {}
inline GLint glGetUniformLocation(GLuint, const GLchar*) { return 0; }
inline void glUniform1i(GLint, GLint) {}
inline void glUniform1ui(GLint, GLuint) {}
inline void glUniform3fv(GLint, GLsizei, const GLfloat*) {}
inline void glUniformMatrix4fv(GLint, GLsizei, GLboolean, const GLfloat*) {}
inline void glViewport(GLint, GLint, GLsizei, GLsizei) {}
inline void glClear(GLbitfield) {}
inline void glClearColor(GLfloat, GLfloat, GLfloat, GLfloat) {}
inline void glEnable(GLenum) {}
inline void glDrawArrays(GLenum, GLint, GLsizei) {}
inline GLuint glCreateShader(GLenum) { return ezrt_gl_capture().next_name++; }
inline void glShaderSource(GLuint, GLsizei, const GLchar**, const GLint*) {}
inline void glCompileShader(GLuint) {}
inline void glGetShaderiv(GLuint, GLenum, GLint* p) { *p = 1; }
inline void glGetShaderInfoLog(GLuint, GLsizei, GLsizei*, GLchar* log) { if (log) log[0] = 0; }
inline GLuint glCreateProgram() { return ezrt_gl_capture().next_name++; }
inline void glAttachShader(GLuint, GLuint) {}
inline void glLinkProgram(GLuint) {}
inline void glDeleteShader(GLuint) {}
inline int glewInit() { return 0; }

#endif

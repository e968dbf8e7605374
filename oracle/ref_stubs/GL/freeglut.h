// GL/freeglut.h -- STAND-IN (no window system here).  *** TEST INFRASTRUCTURE (oracle/), NOT PRODUCT.
// glutMainLoop() returns at once: the reference's main() ends after it has built and uploaded the scene.
#ifndef EZRT_STUB_FREEGLUT_H
#define EZRT_STUB_FREEGLUT_H
enum { GLUT_RGBA = 0, GLUT_DEPTH = 16, GLUT_LEFT_BUTTON = 0, GLUT_DOWN = 0 };
inline void glutInit(int*, char**) {}
inline void glutInitDisplayMode(unsigned int) {}
inline void glutInitWindowSize(int, int) {}
inline void glutInitWindowPosition(int, int) {}
inline int glutCreateWindow(const char*) { return 1; }
inline void glutDisplayFunc(void (*)(void)) {}
inline void glutIdleFunc(void (*)(void)) {}
inline void glutMotionFunc(void (*)(int, int)) {}
inline void glutMouseFunc(void (*)(int, int, int, int)) {}
inline void glutMouseWheelFunc(void (*)(int, int, int, int)) {}
inline void glutMainLoop() {}
inline void glutSwapBuffers() {}
inline void glutPostRedisplay() {}
#endif

/* Headless stand-in for include/gl_helper.h (GL/GLUT headers are absent on the B200 boxes). */
#ifndef MORT_ORACLE_STUB_GL_HELPER_H
#define MORT_ORACLE_STUB_GL_HELPER_H
#endif

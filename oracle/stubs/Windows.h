/* Headless stand-in for <Windows.h> so the UNMODIFIED reference translation unit
 * (/root/reference/mort.cu:5,52-75,736-737) compiles on Linux.  Test infrastructure only:
 * nothing here is reachable from the product library.  The reference's input() /
 * main() that use these symbols are compiled but never called by the harness. */
#ifndef MORT_ORACLE_STUB_WINDOWS_H
#define MORT_ORACLE_STUB_WINDOWS_H
typedef struct tagPOINT { long x; long y; } POINT;
#define VK_LBUTTON 0x01
static inline short GetKeyState(int) { return 0; }
static inline int GetCursorPos(POINT* p) { p->x = 0; p->y = 0; return 1; }
#endif

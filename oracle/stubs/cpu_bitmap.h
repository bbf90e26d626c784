/* Headless stand-in for include/cpu_bitmap.h (unused by the render path). */
#ifndef MORT_ORACLE_STUB_CPU_BITMAP_H
#define MORT_ORACLE_STUB_CPU_BITMAP_H
#endif

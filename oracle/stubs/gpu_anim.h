/* Headless stand-in for include/gpu_anim.h (GLUT + GL interop, /root/reference/include/gpu_anim.h:32-110).
 * Only the surface mort.cu:729-743 touches is declared; there is no window and no GL. */
#ifndef MORT_ORACLE_STUB_GPU_ANIM_H
#define MORT_ORACLE_STUB_GPU_ANIM_H
#include <cuda_runtime.h>
struct GPUAnimBitmap {
    int width, height;
    void* dataBlock;
    GPUAnimBitmap(int w, int h, void* d = 0) : width(w), height(h), dataBlock(d) {}
    void anim_and_exit(void (*)(uchar4*, void*, int), void (*)(void*)) {}
};
#endif

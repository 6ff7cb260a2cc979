// Host build of csrc/boxgeom.cuh for CPU-side verification of the device geometry code (tests only).
#include "../../marie-icr_b200/csrc/boxgeom.cuh"
#include <stdlib.h>
extern "C" int host_component_box(const short* rowmin, const short* rowmax, int y0, int h, int sx, int ex, int sy,
                                  int ey, int niter, float* box) {
    MbHullWork* w = (MbHullWork*)malloc(sizeof(MbHullWork));
    int rc = mb_component_box(w, rowmin, rowmax, y0, h, sx, ex, sy, ey, niter, box);
    free(w);
    return rc;
}
extern "C" void host_adjust_and_rect(const float* box, double sx, double sy, int img_w, int img_h, float* adj,
                                     int* rect) {
    mb_adjust_and_rect(box, sx, sy, img_w, img_h, adj, rect);
}
// the compact workspace of the one-thread-per-box kernel (boxes.cu box_extract_small_kernel); -1 = overflow
extern "C" int host_component_box_small(const short* rowmin, const short* rowmax, int y0, int h, int sx, int ex, int sy,
                                        int ey, int niter, float* box) {
    MbHullWorkT<40, MbPtS, short> w;
    return mb_component_box(&w, rowmin, rowmax, y0, h, sx, ex, sy, ey, niter, box);
}

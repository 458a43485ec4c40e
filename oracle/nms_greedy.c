/* Oracle (TEST INFRASTRUCTURE): sequential greedy grid NMS in plain C.
 *
 * Restates corners_nms (reference python/src/nms.py:4-53): candidates are visited in order of
 * descending confidence; a candidate that is still alive is kept and kills every candidate in
 * its (2r+1)x(2r+1) window (nms.py:37-44); a killed candidate kills nothing.  The reference
 * pads its grid by r so border candidates take part (nms.py:33-35); here the window is clamped
 * to the image instead, which is the same thing.
 *
 * The reference orders equal confidences by whatever numpy's unstable argsort returns
 * (nms.py:17), so ties are an allowed difference class; this oracle breaks ties by ascending
 * linear pixel index (y*w + x), and the CUDA path does the same.
 *
 * Build: gcc -O2 -shared -fPIC -o _build/liboracle_nms.so nms_greedy.c   (oracle/build.py)
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { float conf; int32_t idx; } cand_t;

static int by_conf_desc(const void* a, const void* b) {
    const cand_t* p = (const cand_t*)a;
    const cand_t* q = (const cand_t*)b;
    if (p->conf > q->conf) return -1;
    if (p->conf < q->conf) return 1;
    return (p->idx > q->idx) - (p->idx < q->idx);
}

/* heat: h*w fp32 heatmap.  Candidates are heat >= thresh (python/src/netutils.py:59).
 * Writes up to cap survivors, sorted by descending confidence, ties by ascending index, as
 * (x, y, conf); survivors closer than `border` pixels to an image edge are dropped AFTER
 * suppression (python/src/netutils.py:94-99), so they still suppress their neighbours.
 * Returns the number of survivors (which may exceed cap; only cap are written), or -1. */
int oracle_nms_greedy(const float* heat, int h, int w, float thresh, int radius, int border,
                      int cap, int32_t* out_x, int32_t* out_y, float* out_conf) {
    int n = 0, total = h * w, kept = 0;
    for (int i = 0; i < total; ++i) n += heat[i] >= thresh;
    if (n == 0) return 0;
    cand_t* c = (cand_t*)malloc(sizeof(cand_t) * (size_t)n);
    uint8_t* alive = (uint8_t*)calloc((size_t)total, 1);
    if (!c || !alive) { free(c); free(alive); return -1; }
    n = 0;
    for (int i = 0; i < total; ++i)
        if (heat[i] >= thresh) { c[n].conf = heat[i]; c[n].idx = i; alive[i] = 1; ++n; }
    qsort(c, (size_t)n, sizeof(cand_t), by_conf_desc);
    for (int k = 0; k < n; ++k) {
        int i = c[k].idx;
        if (!alive[i]) continue;
        int y = i / w, x = i % w;
        int y0 = y - radius < 0 ? 0 : y - radius, y1 = y + radius >= h ? h - 1 : y + radius;
        int x0 = x - radius < 0 ? 0 : x - radius, x1 = x + radius >= w ? w - 1 : x + radius;
        for (int yy = y0; yy <= y1; ++yy) memset(alive + (size_t)yy * w + x0, 0, (size_t)(x1 - x0 + 1));
        if (x < border || x >= w - border || y < border || y >= h - border) continue;
        if (kept < cap) { out_x[kept] = x; out_y[kept] = y; out_conf[kept] = c[k].conf; }
        ++kept;
    }
    free(c); free(alive);
    return kept;
}

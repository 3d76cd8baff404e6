/* TEST INFRASTRUCTURE ONLY — plain-C restatement of the reference's CPU NMS routines.
 *
 *   oracle_cpu_nms       follows utils/nms/cpu_nms.pyx:17-68  (suppress_on_equal=1: "ovr >= thresh")
 *                        and  utils/nms/nms_kernel.cu:24-32,71 + :124-140 / py_cpu_nms.py:10-38
 *                        (suppress_on_equal=0: "ovr > thresh").  +1 pixel area convention.
 *   oracle_cpu_soft_nms  follows utils/nms/cpu_nms.pyx:70-163 (in-place, returns N_final).
 *
 * Candidate order: score descending, original index ascending on ties (the reference uses an
 * unstable argsort, so tie order is unspecified upstream).  All arithmetic is float (fp32).
 * Built by oracle/build.py with -O2 -ffp-contract=off so no FMA contraction changes rounding.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct { float s; int i; } key_t_;

static int cmp_desc(const void* a, const void* b) {
  const key_t_* x = (const key_t_*)a; const key_t_* y = (const key_t_*)b;
  if (x->s > y->s) return -1;
  if (x->s < y->s) return 1;
  return (x->i > y->i) - (x->i < y->i);
}

static inline float fmaxf_(float a, float b) { return a >= b ? a : b; }
static inline float fminf_(float a, float b) { return a <= b ? a : b; }

int oracle_cpu_nms(const float* dets, int n, float thresh, int suppress_on_equal, int* keep) {
  if (n <= 0) return 0;
  key_t_* order = (key_t_*)malloc(sizeof(key_t_) * (size_t)n);
  float* areas = (float*)malloc(sizeof(float) * (size_t)n);
  unsigned char* suppressed = (unsigned char*)calloc((size_t)n, 1);
  for (int i = 0; i < n; ++i) {
    const float* d = dets + 5 * (size_t)i;
    areas[i] = (d[2] - d[0] + 1) * (d[3] - d[1] + 1);
    order[i].s = d[4]; order[i].i = i;
  }
  qsort(order, (size_t)n, sizeof(key_t_), cmp_desc);
  int nk = 0;
  for (int _i = 0; _i < n; ++_i) {
    int i = order[_i].i;
    if (suppressed[i]) continue;
    keep[nk++] = i;
    const float* a = dets + 5 * (size_t)i;
    float iarea = areas[i];
    for (int _j = _i + 1; _j < n; ++_j) {
      int j = order[_j].i;
      if (suppressed[j]) continue;
      const float* b = dets + 5 * (size_t)j;
      float xx1 = fmaxf_(a[0], b[0]), yy1 = fmaxf_(a[1], b[1]);
      float xx2 = fminf_(a[2], b[2]), yy2 = fminf_(a[3], b[3]);
      float w = fmaxf_(0.0f, xx2 - xx1 + 1), h = fmaxf_(0.0f, yy2 - yy1 + 1);
      float inter = w * h;
      float ovr = inter / (iarea + areas[j] - inter);
      if (suppress_on_equal ? (ovr >= thresh) : (ovr > thresh)) suppressed[j] = 1;
    }
  }
  free(order); free(areas); free(suppressed);
  return nk;
}

int oracle_cpu_soft_nms(float* boxes, int n, float sigma, float Nt, float threshold, unsigned method) {
  int N = n;
  for (int i = 0; i < N; ++i) {
    float maxscore = boxes[i * 5 + 4];
    int maxpos = i;
    float t[5];
    memcpy(t, boxes + i * 5, sizeof t);
    for (int pos = i + 1; pos < N; ++pos)
      if (maxscore < boxes[pos * 5 + 4]) { maxscore = boxes[pos * 5 + 4]; maxpos = pos; }
    memcpy(boxes + i * 5, boxes + maxpos * 5, sizeof t);
    memcpy(boxes + maxpos * 5, t, sizeof t);
    float tx1 = boxes[i * 5], ty1 = boxes[i * 5 + 1], tx2 = boxes[i * 5 + 2], ty2 = boxes[i * 5 + 3];
    int pos = i + 1;
    while (pos < N) {
      float* b = boxes + pos * 5;
      float x1 = b[0], y1 = b[1], x2 = b[2], y2 = b[3];
      /* NB: Cython emits the literal 1 as the double constant 1.0, so these sub-expressions are
       * evaluated in double and rounded to float once, on assignment (generated cpu_nms.c; the
       * float-only evaluation differs by 1 ulp on ~5% of pairs). */
      float area = (float)(((double)(x2 - x1) + 1.0) * ((double)(y2 - y1) + 1.0));
      float iw = (float)((double)(fminf_(tx2, x2) - fmaxf_(tx1, x1)) + 1.0);
      if (iw > 0) {
        float ih = (float)((double)(fminf_(ty2, y2) - fmaxf_(ty1, y1)) + 1.0);
        if (ih > 0) {
          float ua = (float)(((((double)(tx2 - tx1) + 1.0) * ((double)(ty2 - ty1) + 1.0)) + (double)area)
                             - (double)(iw * ih));
          float ov = (iw * ih) / ua;
          float weight;
          if (method == 1) weight = ov > Nt ? 1 - ov : 1;
          else if (method == 2) weight = (float)exp((double)(-(ov * ov) / sigma));
          else weight = ov > Nt ? 0 : 1;
          b[4] = weight * b[4];
          if (b[4] < threshold) {
            memcpy(b, boxes + (N - 1) * 5, sizeof t);
            N -= 1;
            pos -= 1;
          }
        }
      }
      pos += 1;
    }
  }
  return N;
}

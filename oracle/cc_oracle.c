/* TEST INFRASTRUCTURE — CPU restatement (plain C, sequential) of the reference's connected-component
 * labelling and hole filling.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg
 * may call this; the product never does.
 *
 * Follows /root/reference/sam2/csrc/connected_components.cu:
 *   - 8-connectivity between non-zero pixels of a uint8 [N,1,H,W] mask          (:72-118 merge)
 *   - label of a component = 1 + min over its pixels of the raster index of the pixel's 2x2 block
 *     origin ((row & ~1) * W + (col & ~1)); background label 0                   (:62-70, :129-168)
 *   - counts[p] = area of p's component, 0 on background                          (:170-209)
 * and /root/reference/sam2/utils/misc.py:365-393 (fill_holes_in_mask_scores):
 *   holes = components of (score <= 0); pixels in components of area <= max_area get score 0.1.
 *
 * Parity note: the reference kernel itself cannot run in the build container (no GPU) nor on the GPU
 * box (no /root/reference there), so the label *numbering* convention above is pinned by reading the
 * source, not by execution; component membership and areas are convention-free.
 */
#include <stdint.h>
#include <stdlib.h>
#include <limits.h>

static void label_one(const uint8_t* m, int32_t* labels, int32_t* counts, int H, int W, int32_t* stack) {
  const int HW = H * W;
  for (int i = 0; i < HW; ++i) { labels[i] = 0; counts[i] = 0; }
  /* labels used as "visited" marker (-1) during the flood fill */
  for (int s = 0; s < HW; ++s) {
    if (!m[s] || labels[s] != 0) continue;
    int top = 0, n = 0;
    int32_t minblk = INT_MAX;
    stack[top++] = s;
    labels[s] = -1;
    /* first pass: collect component (members are kept in the tail of `stack` array) */
    int head = 0;
    while (head < top) {
      const int p = stack[head++];
      const int y = p / W, x = p % W;
      const int32_t blk = (y & ~1) * W + (x & ~1);
      if (blk < minblk) minblk = blk;
      ++n;
      for (int dy = -1; dy <= 1; ++dy)
        for (int dx = -1; dx <= 1; ++dx) {
          if (!dy && !dx) continue;
          const int yy = y + dy, xx = x + dx;
          if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
          const int q = yy * W + xx;
          if (m[q] && labels[q] == 0) { labels[q] = -1; stack[top++] = q; }
        }
    }
    for (int i = 0; i < top; ++i) { labels[stack[i]] = minblk + 1; counts[stack[i]] = n; }
  }
}

int cc_oracle_label(const uint8_t* mask, int32_t* labels, int32_t* counts, int N, int H, int W) {
  int32_t* stack = (int32_t*)malloc(sizeof(int32_t) * (size_t)H * W);
  if (!stack) return -1;
  for (int n = 0; n < N; ++n)
    label_one(mask + (size_t)n * H * W, labels + (size_t)n * H * W, counts + (size_t)n * H * W, H, W, stack);
  free(stack);
  return 0;
}

int cc_oracle_fill_holes(float* scores, int N, int H, int W, int max_area) {
  const size_t HW = (size_t)H * W;
  uint8_t* m = (uint8_t*)malloc(HW);
  int32_t* labels = (int32_t*)malloc(sizeof(int32_t) * HW);
  int32_t* counts = (int32_t*)malloc(sizeof(int32_t) * HW);
  int32_t* stack = (int32_t*)malloc(sizeof(int32_t) * HW);
  if (!m || !labels || !counts || !stack) return -1;
  for (int n = 0; n < N; ++n) {
    float* s = scores + (size_t)n * HW;
    for (size_t i = 0; i < HW; ++i) m[i] = s[i] <= 0.0f;
    label_one(m, labels, counts, H, W, stack);
    for (size_t i = 0; i < HW; ++i)
      if (labels[i] > 0 && counts[i] <= max_area) s[i] = 0.1f;
  }
  free(m); free(labels); free(counts); free(stack);
  return 0;
}

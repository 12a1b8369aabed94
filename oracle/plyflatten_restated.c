/* ORACLE (test infrastructure only — never linked into or called by the product path).
 *
 * CPU restatement of the point-cloud rasteriser the reference calls at
 * datasets/satellite_rgb_dep.py:680
 *     dsm = plyflatten(cloud, xoff, yoff, resolution, xsize, ysize, radius=1, sigma=float("inf"))
 * `plyflatten` is a third-party dependency (requirements.txt:11, plyflatten==0.2.0, a ctypes wrapper around
 * `rasterize_cloud` of its C library); it is NOT vendored under /root/reference and not installed in this image, so this
 * file restates the published algorithm of that version:
 *   for every point (x, y, v...) in input order:
 *       i = floor((x - xoff) / resolution);  j = floor((-y - (-yoff)) / resolution)
 *       for k1, k2 in [-radius, radius]:  (ii, jj) = (i + k1, j + k2)
 *           dist = hypotf(x - (xoff + resolution (ii + 0.5)), y - (yoff - resolution (jj + 0.5)))     (float)
 *           weight = sigma == inf ? 1 : exp(-dist^2 / (2 sigma^2))
 *           if (ii, jj) inside the raster:  avg = (v weight + cnt avg) / (weight + cnt);  cnt += weight      (float)
 *   cells with cnt == 0 become NaN.
 * PARITY UNPINNED for this function: there is no plyflatten binary or golden raster in the reference tree to check it
 * against; the reference-side anchors are its call site (arguments, raster shape (ysize, xsize, 1), float32) and the grid
 * derivation at satellite_rgb_dep.py:664-671, which oracle/dsm_np.py restates and pins.
 *
 * Build: gcc -O2 -shared -fPIC -o oracle/_build/libplyflatten_restated.so oracle/plyflatten_restated.c -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

static float distance_weight(float sigma, float d) {
  if (isinf(sigma)) return 1.0f;
  return expf(-d * d / (2.0f * sigma * sigma));
}

/* cloud: nb_points rows of (2 + nb_extra) doubles [x, y, extra...]; raster: ysize*xsize*nb_extra floats (row-major
 * (ysize, xsize, nb_extra)); cnt_out: ysize*xsize floats or NULL.  Returns 0, or -1 on allocation failure. */
int plyflatten_restated(const double* cloud, int64_t nb_points, int nb_extra, double xoff, double yoff, double resolution,
                        int xsize, int ysize, int radius, float sigma, float* raster, float* cnt_out) {
  const int64_t ncell = (int64_t)xsize * ysize;
  float* cnt = (float*)calloc((size_t)ncell, sizeof(float));
  if (!cnt) return -1;
  for (int64_t c = 0; c < ncell * nb_extra; ++c) raster[c] = 0.0f;
  const int stride = 2 + nb_extra;
  for (int64_t p = 0; p < nb_points; ++p) {
    const double xx = cloud[p * stride], yy = cloud[p * stride + 1];
    const int i = (int)floor((xx - xoff) / resolution);
    const int j = (int)floor((-yy - (-yoff)) / resolution);
    for (int k1 = -radius; k1 <= radius; ++k1)
      for (int k2 = -radius; k2 <= radius; ++k2) {
        const int ii = i + k1, jj = j + k2;
        const float dist_x = (float)(xx - (xoff + resolution * (0.5 + ii)));
        const float dist_y = (float)(yy - (yoff - resolution * (0.5 + jj)));
        const float dist = hypotf(dist_x, dist_y);
        const float weight = distance_weight(sigma, dist);
        if (ii < 0 || jj < 0 || ii >= xsize || jj >= ysize) continue;
        const int64_t k = (int64_t)xsize * jj + ii;
        for (int e = 0; e < nb_extra; ++e) {
          const float v = (float)cloud[p * stride + 2 + e];
          float* avg = &raster[k * nb_extra + e];
          *avg = (v * weight + cnt[k] * *avg) / (weight + cnt[k]);
        }
        cnt[k] += weight;
      }
  }
  for (int64_t k = 0; k < ncell; ++k) {
    if (cnt[k] == 0.0f)
      for (int e = 0; e < nb_extra; ++e) raster[k * nb_extra + e] = NAN;
    if (cnt_out) cnt_out[k] = cnt[k];
  }
  free(cnt);
  return 0;
}

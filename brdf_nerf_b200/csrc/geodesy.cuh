// Geodesy shared by the ray feed (georays.cu) and the DSM point cloud (dsm.cu).  Float64, compiled with -fmad=false.
//   to_ecef      sat_utils.latlon_to_ecef_custom   sat_utils.py:110-125
//   from_ecef    sat_utils.ecef_to_latlon_custom   sat_utils.py:127-146
//   to_utm       sat_utils.utm_from_latlon         sat_utils.py:148-162 -> pyproj "+proj=utm +zone=N" (third-party): Krueger
//                series transverse Mercator to order n^6 (Karney 2011, eqs 7-11, 35), GRS80 (PROJ's default ellipsoid for a
//                +proj string that names none), k0 = 0.9996, false easting 500 km, and no false northing: the reference
//                builds "+zone=<n><letter>", which carries no "+south".  Restated in oracle/georays_np.py.
#pragma once
#include <math.h>

namespace bn {

constexpr double kPi = 3.141592653589793;

struct UtmParams {
  double lon0_deg, A, e;       // central meridian, k0 * rectifying radius, eccentricity
  double alpha[6];             // Krueger series coefficients
};

inline UtmParams make_utm_params(int zone) {
  UtmParams u;
  const double f = 1.0 / 298.257222101, n = f / (2 - f);
  const double n2 = n * n, n3 = n2 * n, n4 = n3 * n, n5 = n4 * n, n6 = n5 * n;
  u.e = sqrt(f * (2 - f));
  u.A = 0.9996 * (6378137.0 / (1 + n) * (1 + n2 / 4 + n4 / 64 + n6 / 256));
  u.lon0_deg = (double)((zone - 1) * 6 - 180 + 3);
  u.alpha[0] = n / 2 - 2 * n2 / 3 + 5 * n3 / 16 + 41 * n4 / 180 - 127 * n5 / 288 + 7891 * n6 / 37800;
  u.alpha[1] = 13 * n2 / 48 - 3 * n3 / 5 + 557 * n4 / 1440 + 281 * n5 / 630 - 1983433 * n6 / 1935360;
  u.alpha[2] = 61 * n3 / 240 - 103 * n4 / 140 + 15061 * n5 / 26880 + 167603 * n6 / 181440;
  u.alpha[3] = 49561 * n4 / 161280 - 179 * n5 / 168 + 6601661 * n6 / 7257600;
  u.alpha[4] = 34729 * n5 / 80640 - 3418889 * n6 / 1995840;
  u.alpha[5] = 212378941 * n6 / 319334400;
  return u;
}

__device__ inline void to_ecef(double lat, double lon, double alt, double& x, double& y, double& z) {
  const double rad_lat = lat * (kPi / 180.0), rad_lon = lon * (kPi / 180.0);
  const double a = 6378137.0, finv = 298.257223563, f = 1 / finv, e2 = 1 - (1 - f) * (1 - f);
  const double sl = sin(rad_lat), cl = cos(rad_lat);
  const double v = a / sqrt(1 - e2 * sl * sl);
  x = (v + alt) * cl * cos(rad_lon);
  y = (v + alt) * cl * sin(rad_lon);
  z = (v * (1 - e2) + alt) * sl;
}

__device__ inline void from_ecef(double x, double y, double z, double& lat, double& lon, double& alt) {
  const double a = 6378137.0, e = 8.1819190842622e-2;
  const double asq = a * a, esq = e * e;
  const double b = sqrt(asq * (1 - esq)), bsq = b * b;
  const double ep = sqrt((asq - bsq) / bsq);
  const double p = sqrt((x * x) + (y * y));
  const double th = atan2(a * z, b * p);
  const double st = sin(th), ct = cos(th);
  const double lon_r = atan2(y, x);
  const double lat_r = atan2((z + (ep * ep) * b * (st * st * st)), (p - esq * a * (ct * ct * ct)));
  const double sl = sin(lat_r);
  const double N = a / (sqrt(1 - esq * (sl * sl)));
  alt = p / cos(lat_r) - N;
  lon = lon_r * 180 / kPi;
  lat = lat_r * 180 / kPi;
}

__device__ inline void to_utm(const UtmParams& u, double lat, double lon, double& east, double& north) {
  const double phi = lat * (kPi / 180.0), lam = (lon - u.lon0_deg) * (kPi / 180.0);
  const double tau = tan(phi);
  const double sigma = sinh(u.e * atanh(u.e * tau / sqrt(1 + tau * tau)));
  const double taup = tau * sqrt(1 + sigma * sigma) - sigma * sqrt(1 + tau * tau);
  const double cl = cos(lam);
  const double xip = atan2(taup, cl);
  const double etap = asinh(sin(lam) / sqrt(taup * taup + cl * cl));
  double xi = xip, eta = etap;
#pragma unroll
  for (int j = 1; j <= 6; ++j) {
    xi = xi + u.alpha[j - 1] * sin(2 * j * xip) * cosh(2 * j * etap);
    eta = eta + u.alpha[j - 1] * cos(2 * j * xip) * sinh(2 * j * etap);
  }
  east = 500000.0 + u.A * eta;
  north = u.A * xi;
}

}  // namespace bn

// Geometry formulas of the reference's set-up, shared by the general builder (mesh_build.cpp)
// and the structured per-rank generator (structured.cpp) so that both produce the same bits:
//   face area vector + centroid   src/setup/calc_aip_xyzip.f90:25-72  (triangle fan from vertex 1)
//   cell volume + centroid        src/setup/calc_vol_cv_centers.f90:20-55 (face pyramids about the vertex mean)
#pragma once

#if defined(__CUDACC__) || defined(CUEMU)
#define CFDL_HD __host__ __device__
#else
#define CFDL_HD
#endif

namespace cfdl {

CFDL_HD inline void cross3(double* a, const double* b, const double* c) {
  a[0] = b[1] * c[2] - b[2] * c[1];
  a[1] = b[2] * c[0] - b[0] * c[2];
  a[2] = b[0] * c[1] - b[1] * c[0];
}

// nl = 3 or 4 vertices r[0..nl) in the owner's face order
CFDL_HD inline void face_area_centroid(const double (*r)[3], int nl, double* aip, double* rip) {
  double dr1[3], dr2[3], areavec[3], subcntr[3], sumcntr[3][3], A[3];
  for (int i = 0; i < 3; ++i) { dr1[i] = r[1][i] - r[0][i]; dr2[i] = r[2][i] - r[0][i]; }
  cross3(areavec, dr1, dr2);
  for (int i = 0; i < 3; ++i) { areavec[i] = 0.5 * areavec[i]; A[i] = areavec[i]; }
  for (int i = 0; i < 3; ++i) subcntr[i] = (r[0][i] + r[1][i] + r[2][i]) / 3.0;
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) sumcntr[i][j] = subcntr[i] * areavec[j];
  if (nl == 4) {
    for (int i = 0; i < 3; ++i) { dr1[i] = r[2][i] - r[0][i]; dr2[i] = r[3][i] - r[0][i]; }
    cross3(areavec, dr1, dr2);
    for (int i = 0; i < 3; ++i) { areavec[i] = 0.5 * areavec[i]; A[i] = A[i] + areavec[i]; }
    for (int i = 0; i < 3; ++i) subcntr[i] = (r[0][i] + r[2][i] + r[3][i]) / 3.0;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) sumcntr[i][j] = sumcntr[i][j] + subcntr[i] * areavec[j];
  }
  const double aa = A[0] * A[0] + A[1] * A[1] + A[2] * A[2];
  for (int i = 0; i < 3; ++i) {
    double c = 0.0;
    for (int j = 0; j < 3; ++j) c = c + sumcntr[i][j] * A[j];
    aip[i] = A[i];
    rip[i] = c / aa;
  }
}

// running sums of the pyramid decomposition of one cell; gc = mean of the cell's vertices
struct CellAccumulator {
  double gc[3], sum_vol = 0.0, rc[3] = {0.0, 0.0, 0.0};
  CFDL_HD void add_face(int sg, const double* aip, const double* rip) {
    const double h[3] = {rip[0] - gc[0], rip[1] - gc[1], rip[2] - gc[2]};
    const double sub_vol = ((sg * aip[0]) * h[0] + (sg * aip[1]) * h[1] + (sg * aip[2]) * h[2]) / 3.0;
    sum_vol = sum_vol + sub_vol;
    for (int i = 0; i < 3; ++i) rc[i] = rc[i] + (0.25 * gc[i] + 0.75 * rip[i]) * sub_vol;
  }
  CFDL_HD void finish(double* centre, double* vol) const {
    centre[0] = rc[0] / sum_vol; centre[1] = rc[1] / sum_vol; centre[2] = rc[2] / sum_vol;
    *vol = sum_vol;
  }
};

}  // namespace cfdl

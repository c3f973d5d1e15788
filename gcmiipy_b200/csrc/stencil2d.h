// stencil2d.h -- neighbourhood addressing shared by the 2-D schemes (reference coordinates.py:29-79).
//
// A neighbourhood holds the linear row offsets of rows j-1 .. j+2 and the column indices i-1 .. i+1 of one
// cell.  On global arrays the entries wrap periodically exactly like np.roll (constants.py:85); on a
// shared-memory tile they are plain tile offsets (the wrap was resolved when the tile was filled).
#pragma once
#include "gcm_common.h"

struct GcmNb {
  int row[4];  // row offset (already multiplied by the pitch) of j-1, j, j+1, j+2
  int col[3];  // i-1, i, i+1
  __device__ __forceinline__ double operator()(const double* a, int dj, int di) const {
    return a[row[dj + 1] + col[di + 1]];
  }
};

__device__ __forceinline__ GcmNb gcm_nb_global(int j, int i, int H, int W) {
  GcmNb n;
  const int jm = j == 0 ? H - 1 : j - 1;
  const int jp = j + 1 == H ? 0 : j + 1;
  const int jpp = jp + 1 == H ? 0 : jp + 1;
  n.row[0] = jm * W; n.row[1] = j * W; n.row[2] = jp * W; n.row[3] = jpp * W;
  n.col[0] = gcm_im(i, W); n.col[1] = i; n.col[2] = gcm_ip(i, W);
  return n;
}

__device__ __forceinline__ GcmNb gcm_nb_tile(int tj, int ti, int pitch) {
  GcmNb n;
  n.row[0] = (tj - 1) * pitch; n.row[1] = tj * pitch; n.row[2] = (tj + 1) * pitch; n.row[3] = (tj + 2) * pitch;
  n.col[0] = ti - 1; n.col[1] = ti; n.col[2] = ti + 1;
  return n;
}

__host__ __device__ __forceinline__ int gcm_wrap(int x, int n) {
  x %= n;
  return x < 0 ? x + n : x;
}

// pk_mu.h -- the `mu` diagnostic (find_mus, /root/reference/peakondrake/outputting.py:260-370) for one
// realisation, as a serial routine on arrays that is compiled for the device (pk_mu_kernel, one thread per
// realisation walks shared-memory copies of u and du) and for the CPU harness of the not-gpu tests
// (tests/emu/pk_emu.cpp), where it is pinned bit-for-bit against the reference helper's golden vectors.
//
// The reference algorithm is sequential Python with quirks that are kept as written:
//   * a "peak" is a + -> - sign change of the DG0 slope between cells i-1 and i whose CG1 value at the mid-point of
//     the two cell centres exceeds twice the mean of u (:279-286); for i = 0 that mid-point is the middle of the
//     domain because x[-1] is the last cell centre;
//   * np.mean is NumPy's pairwise summation; reproduced here so that the comparison is the same in the last bit;
//   * several peaks: zones between mid-way cuts, only the four highest peaks, the wrap-around cases of :307-326
//     including :316's comparison of a dof index with a peak count.
#pragma once
#include "pk_params.h"

namespace pk {

enum { MU_MAX_PEAKS = 48 };

PK_HD double nan_value() {
#if defined(__CUDA_ARCH__)
  return __longlong_as_double(0x7ff8000000000000LL);
#else
  return std::nan("");
#endif
}

// NumPy's pairwise summation of a contiguous double array (numpy/core/src/umath/loops_utils.h.src, pairwise_sum):
// blocks of at most 128 elements are summed with eight interleaved accumulators, longer ranges are split in two
// (the left half a multiple of 8) and the halves added.  Written without recursion (explicit frames).
PK_HD double np_pairwise_block(const double* a, int n) {
  if (n < 8) {
    double res = 0.0;
    for (int i = 0; i < n; ++i) res += a[i];
    return res;
  }
  double r0 = a[0], r1 = a[1], r2 = a[2], r3 = a[3], r4 = a[4], r5 = a[5], r6 = a[6], r7 = a[7];
  int i;
  for (i = 8; i < n - (n % 8); i += 8) {
    r0 += a[i + 0]; r1 += a[i + 1]; r2 += a[i + 2]; r3 += a[i + 3];
    r4 += a[i + 4]; r5 += a[i + 5]; r6 += a[i + 6]; r7 += a[i + 7];
  }
  double res = ((r0 + r1) + (r2 + r3)) + ((r4 + r5) + (r6 + r7));
  for (; i < n; ++i) res += a[i];
  return res;
}
PK_HD double np_pairwise_sum(const double* a, int n) {
  int lo[32], len[32], stage[32], sp = 0;
  double left[32];
  lo[0] = 0; len[0] = n; stage[0] = 0; left[0] = 0.0;
  double ret = 0.0;
  while (sp >= 0) {
    if (len[sp] <= 128) { ret = np_pairwise_block(a + lo[sp], len[sp]); --sp; continue; }
    int n2 = len[sp] / 2; n2 -= n2 % 8;
    if (stage[sp] == 0) {
      stage[sp] = 1;
      lo[sp + 1] = lo[sp]; len[sp + 1] = n2; stage[sp + 1] = 0; ++sp;
    } else if (stage[sp] == 1) {
      left[sp] = ret; stage[sp] = 2;
      lo[sp + 1] = lo[sp] + n2; len[sp + 1] = len[sp] - n2; stage[sp + 1] = 0; ++sp;
    } else {
      ret = left[sp] + ret; --sp;
    }
  }
  return ret;
}

PK_HD int mu_argext(const double* d, int lo, int hi, bool want_max) {     // first occurrence, like np.argmax/argmin
  int b = lo;
  for (int i = lo + 1; i < hi; ++i)
    if (want_max ? (d[i] > d[b]) : (d[i] < d[b])) b = i;
  return b;
}

// old4 / alt4: the first four entries of list_of_mus / alt_list_of_mus (NaN where the lists are shorter, using the
// length of list_of_mus for both, as the reference's dump loop does, outputting.py:136-146).
// returns the number of peaks found; -1 if more than MU_MAX_PEAKS, -2 where the reference itself would raise
// (the caller then runs the host routine for that realisation, which raises the same error).
PK_HD int mu_serial(const double* u, const double* du, const double* x, int n, double Ld, double* old4, double* alt4) {
  const double nan = nan_value();
  for (int j = 0; j < 4; ++j) { old4[j] = nan; alt4[j] = nan; }
  double L = x[0];
  for (int i = 1; i < n; ++i) if (x[i] > L) L = x[i];
  const double halfL = L / 2;
  const double mean2 = 2 * (np_pairwise_sum(u, n) / (double)n);
  int pidx[MU_MAX_PEAKS]; double pmid[MU_MAX_PEAKS], ph[MU_MAX_PEAKS], px[MU_MAX_PEAKS];
  int npk = 0;
  const double hcell = Ld / n;
  for (int i = 0; i < n; ++i) {
    const int im = (i == 0) ? n - 1 : i - 1;
    if (du[im] > 0 && du[i] < 0) {
      const double mid = 0.5 * (x[im] + x[i]);
      // CG1 point evaluation, the way Function.at does it on this mesh
      const double s = mid / hcell;
      const double fl = floor(s);
      int i0 = (int)fl % n; if (i0 < 0) i0 += n;
      const double th = s - fl;
      const double height = (1 - th) * u[i0] + th * u[(i0 + 1) % n];
      if (height > mean2) {
        if (npk == MU_MAX_PEAKS) return -1;
        pidx[npk] = i; px[npk] = x[i]; pmid[npk] = mid; ph[npk] = height; ++npk;
      }
    }
  }
  int nold = 0, nalt = 0;
  double oldv[MU_MAX_PEAKS], altv[MU_MAX_PEAKS];
  if (npk > 1) {
    double ranked[MU_MAX_PEAKS];
    for (int k = 0; k < npk; ++k) {            // insertion sort, descending
      double v = ph[k]; int j = k;
      while (j > 0 && ranked[j - 1] < v) { ranked[j] = ranked[j - 1]; --j; }
      ranked[j] = v;
    }
    int cuts[MU_MAX_PEAKS];
    for (int k = 0; k < npk; ++k) {
      const int ka = (k == 0) ? npk - 1 : k - 1;
      if (px[k] > px[ka]) cuts[k] = (int)floor(0.5 * (double)(pidx[ka] + pidx[k]));
      else cuts[k] = ((int)floor(0.5 * (double)(pidx[ka] + pidx[k] + n))) % n;
    }
    for (int k = 0; k < npk; ++k) {
      if (!(npk < 4 || ph[k] > ranked[3])) continue;
      const int lo = cuts[(k == 0) ? npk - 1 : k - 1], hi = cuts[k];
      int imax, imin;
      if (hi > lo) { imax = mu_argext(du, lo, hi, true); imin = mu_argext(du, lo, hi, false); }
      else if (hi == 0) { imax = mu_argext(du, lo, n, true); imin = mu_argext(du, lo, n, false); }
      else if (lo == npk - 1) {
        // np.argmax(du[0:lo]) + lo: an index relative to 0 shifted by lo (outputting.py:316-318, as written)
        if (lo == 0) return -2;                                   // empty slice: the reference raises here
        imax = mu_argext(du, 0, lo, true) + lo; imin = mu_argext(du, 0, lo, false) + lo;
        if (imax >= n || imin >= n) return -2;                    // index error in the reference
      } else {
        const int a1 = mu_argext(du, 0, hi, true), a2 = mu_argext(du, lo, n, true);
        const int b1 = mu_argext(du, 0, hi, false), b2 = mu_argext(du, lo, n, false);
        imax = du[a1] > du[a2] ? a1 : a2;
        imin = du[b1] < du[b2] ? b1 : b2;
      }
      const double xmax = x[imax], xmin = x[imin];
      const double d1 = fabs(xmin - xmax), d2 = fabs(xmin - pmid[k]);
      oldv[nold++] = d1 < halfL ? d1 : L - d1;
      altv[nalt++] = d2 < halfL ? d2 : L - d2;
    }
  } else {
    const int imax = mu_argext(du, 0, n, true), imin = mu_argext(du, 0, n, false);
    const double xmax = x[imax], xmin = x[imin];
    const double d1 = fabs(xmin - xmax);
    const double w1 = d1 < halfL ? d1 : L - d1;
    oldv[nold++] = w1;
    if (npk > 0) { const double d2 = fabs(xmin - pmid[0]); altv[nalt++] = d2 < halfL ? d2 : L - d2; }
    else altv[nalt++] = w1 / 2;
  }
  for (int j = 0; j < 4; ++j) {
    if (j < nold) { old4[j] = oldv[j]; alt4[j] = altv[j]; }
  }
  return npk;
}

}  // namespace pk

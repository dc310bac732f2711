/* b200_harness.c -- the drop-in boundary exercised the way the Fortran shims exercise it, without Python/ctypes in
 * between: a compiled program that keeps its arrays in static storage laid out like the model's COMMON blocks
 * (model/inc/SIZE.h dimensions fixed at compile time, CG2D.h / GRID.h / DYNVARS.h style arrays), and calls the
 * `_`-suffixed entry points of include/mitgcm_b200.h with every scalar by reference and LOGICALs as 4-byte integers
 * -- what `CALL CG2D_B200( cg2d_b, cg2d_x, ... )` in mitgcm_b200/fortran/cg2d.F compiles to under the usual Fortran
 * calling convention (lower case, trailing underscore, arguments by address).
 *
 * Input (argv[1]) and output (argv[2]) are flat files of doubles written / read by tests/test_c_harness_gpu.py, which
 * compares the output with the CPU oracle.  Build: gcc -DSNX=.. -DSNY=.. -DOLX=.. -DOLY=.. -DNSX=.. -DNSY=.. -DNR=..
 *   b200_harness.c -I include -L mitgcm_b200/lib -lmitgcm_b200
 * LOGICAL .TRUE. is passed as TRUEVAL (argv[3]): 1 as gfortran stores it, -1 as Intel Fortran does. */
#include <stdio.h>
#include <stdlib.h>
#include "mitgcm_b200.h"

#define PX (SNX + 2 * OLX)
#define PY (SNY + 2 * OLY)

/* COMMON /CG2D_I_RS/ (model/inc/CG2D.h:32-42) and the solver vectors of SOLVE_FOR_PRESSURE (SOLVE_FOR_PRESSURE.h) */
static double aW2d[NSY][NSX][PY][PX], aS2d[NSY][NSX][PY][PX], aC2d[NSY][NSX][PY][PX];
static double pW[NSY][NSX][PY][PX], pS[NSY][NSX][PY][PX], pC[NSY][NSX][PY][PX];
static double cg2d_b[NSY][NSX][PY][PX], cg2d_x[NSY][NSX][PY][PX];
/* a tile3d scratch big enough for any GRID.h array (the shim hands over the COMMON arrays themselves) */
static double gridbuf[NSY][NSX][NR + 1][PY][PX];
/* locals of TEMP_INTEGRATE (temp_integrate.F:80-110) and COMMON /DYNVARS_R/ theta, gT */
static double xA[PY][PX], yA[PY][PX], maskUp[PY][PX], uFld[PY][PX], vFld[PY][PX], wFld[PY][PX];
static double uTrans[PY][PX], vTrans[PY][PX], rTrans[PY][PX], rTransKp1[PY][PX], KappaR[PY][PX];
static double fZon[PY][PX], fMer[PY][PX], fVerT[2][PY][PX];
static double theta[NR][PY][PX], thetaAB[NR][PY][PX], gT[NR][PY][PX];
static double diffKr4[NR], deltaTLev[NR];

static FILE *fin, *fout;
static void rd(double *p, size_t n) {
  if (fread(p, sizeof(double), n, fin) != n) { fprintf(stderr, "harness: short read\n"); exit(2); }
}
static double rd1(void) { double v; rd(&v, 1); return v; }
static void wr(const double *p, size_t n) { fwrite(p, sizeof(double), n, fout); }
static void wr1(double v) { wr(&v, 1); }
static void check(const char *what) {
  if (mitgcm_b200_last_error_() != 0) {
    fprintf(stderr, "harness: %s: error %d: %s\n", what, mitgcm_b200_last_error_(), mitgcm_b200_last_error_string());
    exit(3);
  }
}

int main(int argc, char **argv) {
  if (argc < 4) return 1;
  fin = fopen(argv[1], "rb");
  fout = fopen(argv[2], "wb");
  const int TRUEVAL = atoi(argv[3]), FALSEVAL = 0;
  if (!fin || !fout) return 1;
  const size_t n2 = (size_t)NSY * NSX * PY * PX, ns = (size_t)PY * PX;
  int dims[11] = {SNX, SNY, OLX, OLY, NSX, NSY, NR, 1, 1, 0, 0}, device = -1, ierr = 0;
  mitgcm_b200_init_(dims, &device, &ierr);
  if (ierr) { fprintf(stderr, "harness: init failed: %s\n", mitgcm_b200_last_error_string()); return 3; }

  /* GRID.h: records (id, count, data) */
  int nrec = (int)rd1();
  for (int r = 0; r < nrec; r++) {
    int id = (int)rd1();
    size_t n = (size_t)rd1();
    if (n > sizeof(gridbuf) / sizeof(double)) { fprintf(stderr, "harness: record too large\n"); return 2; }
    rd(&gridbuf[0][0][0][0][0], n);
    mitgcm_b200_set_field_(&id, &gridbuf[0][0][0][0][0], &ierr);
    check("set_field (grid)");
  }
  /* PARAMS.h */
  nrec = (int)rd1();
  for (int r = 0; r < nrec; r++) { int id = (int)rd1(); double v = rd1(); mitgcm_b200_set_param_d_(&id, &v, &ierr); check("set_param_d"); }
  nrec = (int)rd1();
  for (int r = 0; r < nrec; r++) { int id = (int)rd1(); int v = (int)rd1(); mitgcm_b200_set_param_i_(&id, &v, &ierr); check("set_param_i"); }

  /* ---- CALL CG2D( cg2d_b, cg2d_x, firstResidual, minResidualSq, lastResidual, numIters, nIterMin, myThid ) ---- */
  rd(&aW2d[0][0][0][0], n2); rd(&aS2d[0][0][0][0], n2); rd(&aC2d[0][0][0][0], n2);
  rd(&pW[0][0][0][0], n2); rd(&pS[0][0][0][0], n2); rd(&pC[0][0][0][0], n2);
  rd(&cg2d_b[0][0][0][0], n2); rd(&cg2d_x[0][0][0][0], n2);
  {
    const int ids[6] = {MG_AW2D, MG_AS2D, MG_AC2D, MG_PW, MG_PS, MG_PC};
    double *arr[6] = {&aW2d[0][0][0][0], &aS2d[0][0][0][0], &aC2d[0][0][0][0], &pW[0][0][0][0], &pS[0][0][0][0], &pC[0][0][0][0]};
    for (int q = 0; q < 6; q++) { mitgcm_b200_set_field_(&ids[q], arr[q], &ierr); check("set_field (CG2D.h)"); }
  }
  int numIters = (int)rd1(), nIterMin = (int)rd1(), myThid = 1;
  double firstResidual = 0, minResidualSq = 0, lastResidual = 0;
  cg2d_b200_(&cg2d_b[0][0][0][0], &cg2d_x[0][0][0][0], &firstResidual, &minResidualSq, &lastResidual, &numIters, &nIterMin, &myThid);
  check("cg2d_b200_");
  wr(&cg2d_b[0][0][0][0], n2); wr(&cg2d_x[0][0][0][0], n2);
  wr1(firstResidual); wr1(minResidualSq); wr1(lastResidual); wr1(numIters); wr1(nIterMin);

  /* ---- CALL GAD_CALC_RHS( bi, bj, iMin, ..., myThid ) for one level (temp_integrate.F:357) ---- */
  int iarg[10];
  for (int q = 0; q < 10; q++) iarg[q] = (int)rd1();      /* bi bj iMin iMax jMin jMax k kM1 kUp kDown */
  rd(&xA[0][0], ns); rd(&yA[0][0], ns); rd(&maskUp[0][0], ns); rd(&uFld[0][0], ns); rd(&vFld[0][0], ns); rd(&wFld[0][0], ns);
  rd(&uTrans[0][0], ns); rd(&vTrans[0][0], ns); rd(&rTrans[0][0], ns); rd(&rTransKp1[0][0], ns);
  double diffKh = rd1(), diffK4 = rd1();
  rd(&KappaR[0][0], ns); rd(diffKr4, NR);
  rd(&theta[0][0][0], (size_t)NR * ns); rd(&thetaAB[0][0][0], (size_t)NR * ns); rd(deltaTLev, NR);
  int trIdentity = (int)rd1(), advScheme = (int)rd1(), vertAdvScheme = (int)rd1();
  int lflag[7];      /* calcAdvection implicitAdvection applyAB_onTracer trUseDiffKr4 trUseGMRedi trUseKPP trUseSmolHack */
  for (int q = 0; q < 7; q++) lflag[q] = rd1() != 0. ? TRUEVAL : FALSEVAL;
  rd(&fVerT[0][0][0], 2 * ns); rd(&gT[0][0][0], (size_t)NR * ns);
  double myTime = 0.;
  int myIter = 0;
  gad_calc_rhs_b200_(&iarg[0], &iarg[1], &iarg[2], &iarg[3], &iarg[4], &iarg[5], &iarg[6], &iarg[7], &iarg[8], &iarg[9],
                     &xA[0][0], &yA[0][0], &maskUp[0][0], &uFld[0][0], &vFld[0][0], &wFld[0][0], &uTrans[0][0], &vTrans[0][0],
                     &rTrans[0][0], &rTransKp1[0][0], &diffKh, &diffK4, &KappaR[0][0], diffKr4, &theta[0][0][0],
                     &thetaAB[0][0][0], deltaTLev, &trIdentity, &advScheme, &vertAdvScheme, &lflag[0], &lflag[1], &lflag[2],
                     &lflag[3], &lflag[4], &lflag[5], &lflag[6], &fZon[0][0], &fMer[0][0], &fVerT[0][0][0], &gT[0][0][0],
                     &myTime, &myIter, &myThid);
  check("gad_calc_rhs_b200_");
  wr(&fZon[0][0], ns); wr(&fMer[0][0], ns); wr(&fVerT[0][0][0], 2 * ns); wr(&gT[0][0][0], (size_t)NR * ns);
  wr1((double)mitgcm_b200_launch_count_());
  mitgcm_b200_finalize_();
  fclose(fin);
  fclose(fout);
  return 0;
}

// The scalar step of the one-sweep Jacobi-PCG at its single reduction point (fsb_cg_one.cu),
// shared by the device kernel and the host emulation in tests/cpu_emul.
// Eigen's loop for comparison: SURVEY.md Appendix B; src/FluidSolver.cpp:418-426.
#pragma once

#if defined(__CUDACC__)
#define FSB_HD __host__ __device__
#else
#define FSB_HD
#endif

// per-CTA copy of the CG scalars: every CTA derives the same values from the same totals
struct OneState
{
  double rz, r2;
  float alpha, beta, thr;
  int iter, done, max_iters, comm_error;
  int sweep; // -1: the set-up sweep (alpha = beta = 0), k >= 0: iteration k
  unsigned long long seq;
  volatile unsigned int released; // last sweep whose reduction this CTA has completed
};

// Totals of the sweep that formed r' = r_{k+1}, z' = D^-1 r', p' = p_{k+1}, q' = A p':
//   pq = p'.q'   rz = r'.z'   r2 = |r'|^2   zq = z'.q'   qmq = q'.D^-1 q'
FSB_HD inline void one_advance(OneState* ss, double pq, double rz, double r2, double zq, double qmq)
{
  if (ss->sweep >= 0)
  {
    ss->r2 = r2;
    if ((float)r2 < ss->thr) ss->done = 1; // converged: Eigen breaks before i++
    else
    {
      ss->iter = ss->iter + 1;
      if (ss->iter >= ss->max_iters) ss->done = 1;
    }
  }
  if (!ss->done)
  {
    ss->rz = rz;
    const float a = (float)rz / (float)pq; // Eigen: alpha = absNew / p.dot(tmp)
    // r''.z'' of the NEXT residual r'' = r' - a q', exact algebra on the stored vectors:
    const double ad = (double)a;
    const double rz_next = rz - 2.0 * ad * zq + ad * ad * qmq;
    ss->alpha = a;
    ss->beta = (float)rz_next / (float)rz; // Eigen: beta = absNew / absOld
  }
  ss->sweep = ss->sweep + 1;
}

// One valence interaction, from term lists to accumulated energy / gradient / virial.  Shared between the CUDA kernel
// (evalence.cu) and the CPU harness of the tests (tests/valmath_host.cpp) through the accumulator policy `Acc`:
//    acc.energy(term, e)   acc.grad(atom, gx, gy, gz)   acc.virial(v6)
// The eight terms are laid out back to back in one index space (off[0..8]) in the order of valparams.TERMS, the
// way the reference fuses them into one launch (src/cu/evalence.cu:17-330).
#pragma once
#include "valmath.cuh"

namespace vm {
enum { T_BOND = 0, T_ANGLE, T_STRBND, T_UREY, T_OPBEND, T_TORSION, T_PITORS, T_TORTOR, T_COUNT };

template <class R>
struct ValDev {
   int off[T_COUNT + 1];      // off[t+1] - off[t] interactions of term t (0 when the term is switched off)
   const int* ibnd;           // [nb][2]
   const R* bprm;             // [nb][2]   force constant, ideal length
   const int* iang;           // [na][4]
   const R* aprm;             // [na][2]   force constant, ideal angle (deg)
   const int* angtyp;         // [na]      0 harmonic, 1 in-plane
   const int* isb;            // [nsb][3]
   const R* sprm;             // [nsb][5]  k1, k2, ideal angle, ideal a-b, ideal c-b
   const int* iury;           // [nu][3]
   const R* uprm;             // [nu][2]
   const int* iopb;           // [nopb][4] a, b (centre), c, d
   const R* oprm;             // [nopb]
   const int* itors;          // [nt][4]
   const R* tprm;             // [nt][18]  6 x {amplitude, cos(phase), sin(phase)}
   const int* ipit;           // [npt][6]
   const R* pprm;             // [npt]
   const int* itt;            // [ntt][5]
   const int* ttchk;          // [ntt]     chirality probe atom or -1
   const int* ttgrid;         // [ntt]
   const TorTorGrid* grids;
   const R *ttx, *tty, *tbf, *tbx, *tby, *tbxy;
   Consts<R> K;
   int opbtyp;                // 0 W-D-C, 1 Allinger
};

template <class R>
VM_HD void load_rel(const double* xyz, const int* ia, int m, V3<R>* X)
{
   const double x0 = xyz[3 * ia[0]], y0 = xyz[3 * ia[0] + 1], z0 = xyz[3 * ia[0] + 2];
   X[0] = mk<R>(0, 0, 0);
   for (int k = 1; k < m; ++k)
      X[k] = mk<R>((R)(xyz[3 * ia[k]] - x0), (R)(xyz[3 * ia[k] + 1] - y0), (R)(xyz[3 * ia[k] + 2] - z0));
}

template <class R, class Acc>
VM_HD void eval_interaction(const ValDev<R>& D, int idx, const double* xyz, bool do_g, bool do_v, Acc& acc)
{
   int term = 0;
   while (term < T_COUNT - 1 && idx >= D.off[term + 1])
      ++term;
   const int i = idx - D.off[term];
   V3<R> X[6], G[6];
   R vir[6];
   R e = 0;
   const int* ia = nullptr;
   int m = 0;
   int ia_local[3];
   bool own_vir = false;
   switch (term) {
   case T_BOND:
      ia = D.ibnd + 2 * i, m = 2;
      load_rel(xyz, ia, m, X);
      e = stretch(X, D.bprm[2 * i + 1], D.bprm[2 * i], D.K.bndunit, D.K.cbnd, D.K.qbnd, G);
      break;
   case T_ANGLE: {
      const bool inpl = D.angtyp[i] == 1;
      ia = D.iang + 4 * i, m = inpl ? 4 : 3;
      load_rel(xyz, ia, m, X);
      e = angle_bend(X, inpl, D.aprm[2 * i + 1], D.aprm[2 * i], D.K, G);
      break;
   }
   case T_STRBND:
      ia = D.isb + 3 * i, m = 3;
      load_rel(xyz, ia, m, X);
      e = stretch_bend(X, D.sprm[5 * i + 2], D.sprm[5 * i + 3], D.sprm[5 * i + 4], D.sprm[5 * i], D.sprm[5 * i + 1], D.K.stbnunit, G);
      break;
   case T_UREY:
      ia_local[0] = D.iury[3 * i], ia_local[1] = D.iury[3 * i + 2];
      ia = ia_local, m = 2;
      load_rel(xyz, ia, m, X);
      e = stretch(X, D.uprm[2 * i + 1], D.uprm[2 * i], D.K.ureyunit, D.K.cury, D.K.qury, G);
      break;
   case T_OPBEND:
      ia = D.iopb + 4 * i, m = 4;
      load_rel(xyz, ia, m, X);
      e = opbend(X, D.opbtyp == 1, D.oprm[i], D.K, G);
      break;
   case T_TORSION:
      ia = D.itors + 4 * i, m = 4;
      load_rel(xyz, ia, m, X);
      e = torsion(X, D.tprm + 18 * i, D.K.torsunit, G);
      break;
   case T_PITORS:
      ia = D.ipit + 6 * i, m = 6;
      load_rel(xyz, ia, m, X);
      e = pitors(X, D.pprm[i], D.K.ptorunit, G, vir);
      own_vir = true;
      break;
   default: {
      ia = D.itt + 5 * i, m = 5;
      load_rel(xyz, ia, m, X);
      const int chk = D.ttchk[i];
      V3<R> pc = mk<R>(0, 0, 0);
      if (chk >= 0)
         pc = mk<R>((R)(xyz[3 * chk] - xyz[3 * ia[0]]), (R)(xyz[3 * chk + 1] - xyz[3 * ia[0] + 1]), (R)(xyz[3 * chk + 2] - xyz[3 * ia[0] + 2]));
      e = tortor(X, chk >= 0, pc, D.grids[D.ttgrid[i]], D.ttx, D.tty, D.tbf, D.tbx, D.tby, D.tbxy, D.K.ttorunit, G);
      break;
   }
   }
   acc.energy(term, e);
   if (do_g) {
      for (int k = 0; k < m; ++k)
         acc.grad(ia[k], G[k].x, G[k].y, G[k].z);
      if (do_v) {
         if (!own_vir)
            virial6(X, G, m, vir);
         acc.virial(vir);
      }
   }
}
}      // namespace vm

// Host-side packing of an apx_valence description into the flat arrays vm::ValDev points at.  Used by evalence.cu
// (which uploads them) and by the CPU harness of the tests (tests/valmath_host.cpp).
#pragma once
#include "apx.h"
#include "valterms.cuh"
#include <cmath>
#include <vector>

template <class R>
struct ValPacked {
   std::vector<int> ibnd, iang, angtyp, isb, iury, iopb, itors, ipit, itt, ttchk, ttgrid;
   std::vector<R> bprm, aprm, sprm, uprm, oprm, tprm, pprm, ttx, tty, tbf, tbx, tby, tbxy;
   std::vector<vm::TorTorGrid> grids;
   vm::Consts<R> K;
   int opbtyp = 0;
   int off[vm::T_COUNT + 1];
   int count[vm::T_COUNT];

   explicit ValPacked(const apx_valence& v)
   {
      const double deg = 3.141592653589793238 / 180.0;
      auto ints = [](const int* p, size_t n) { return std::vector<int>(p, p + n); };
      ibnd = ints(v.ibnd, 2 * (size_t)v.nbond);
      for (int i = 0; i < v.nbond; ++i)
         bprm.push_back((R)v.bk[i]), bprm.push_back((R)v.bl[i]);
      iang = ints(v.iang, 4 * (size_t)v.nangle);
      angtyp = ints(v.angtyp, (size_t)v.nangle);
      for (int i = 0; i < v.nangle; ++i)
         aprm.push_back((R)v.ak[i]), aprm.push_back((R)v.anat[i]);
      isb = ints(v.isb, 3 * (size_t)v.nstrbnd);
      for (int i = 0; i < v.nstrbnd; ++i) {
         sprm.push_back((R)v.sbk[2 * i]), sprm.push_back((R)v.sbk[2 * i + 1]), sprm.push_back((R)v.sb_anat[i]);
         sprm.push_back((R)v.sb_bl[2 * i]), sprm.push_back((R)v.sb_bl[2 * i + 1]);
      }
      iury = ints(v.iury, 3 * (size_t)v.nurey);
      for (int i = 0; i < v.nurey; ++i)
         uprm.push_back((R)v.uk[i]), uprm.push_back((R)v.ul[i]);
      iopb = ints(v.iopb, 4 * (size_t)v.nopbend);
      for (int i = 0; i < v.nopbend; ++i)
         oprm.push_back((R)v.opbk[i]);
      itors = ints(v.itors, 4 * (size_t)v.ntors);
      for (int i = 0; i < v.ntors; ++i)
         for (int k = 0; k < 6; ++k) {      // ktors.f:536-553 keeps cos and sin of every phase
            const double ph = v.tors_phase[6 * i + k] * deg;
            tprm.push_back((R)v.tors_v[6 * i + k]), tprm.push_back((R)std::cos(ph)), tprm.push_back((R)std::sin(ph));
         }
      ipit = ints(v.ipit, 6 * (size_t)v.npitors);
      for (int i = 0; i < v.npitors; ++i)
         pprm.push_back((R)v.kpit[i]);
      itt = ints(v.itt, 5 * (size_t)v.ntortor);
      ttchk = ints(v.tt_chk, (size_t)v.ntortor);
      ttgrid = ints(v.tt_grid, (size_t)v.ntortor);
      size_t nx = 0, ny = 0, nf = 0;
      for (int g = 0; g < v.ngrid; ++g) {
         vm::TorTorGrid t;
         t.nx = v.tnx[g], t.ny = v.tny[g], t.off = v.tt_off[g], t.xoff = v.tt_xoff[g], t.yoff = v.tt_yoff[g];
         grids.push_back(t);
         nx += t.nx, ny += t.ny, nf += (size_t)t.nx * t.ny;
      }
      auto reals = [](const double* p, size_t n) {
         std::vector<R> o(n);
         for (size_t i = 0; i < n; ++i)
            o[i] = (R)p[i];
         return o;
      };
      ttx = reals(v.ttx, nx), tty = reals(v.tty, ny);
      tbf = reals(v.tbf, nf), tbx = reals(v.tbx, nf), tby = reals(v.tby, nf), tbxy = reals(v.tbxy, nf);
      R* k = reinterpret_cast<R*>(&K);
      for (int i = 0; i < 20; ++i)
         k[i] = (R)v.consts[i];
      opbtyp = v.opbtyp;
      const int cnt[vm::T_COUNT] = {v.nbond, v.nangle, v.nstrbnd, v.nurey, v.nopbend, v.ntors, v.npitors, v.ntortor};
      off[0] = 0;
      for (int t = 0; t < vm::T_COUNT; ++t) {
         count[t] = v.use[t] ? cnt[t] : 0;
         off[t + 1] = off[t] + count[t];
      }
   }

   // view over these host vectors (the CUDA side builds the same struct over device copies)
   vm::ValDev<R> view() const
   {
      vm::ValDev<R> D;
      for (int t = 0; t <= vm::T_COUNT; ++t)
         D.off[t] = off[t];
      D.ibnd = ibnd.data(), D.bprm = bprm.data(), D.iang = iang.data(), D.aprm = aprm.data(), D.angtyp = angtyp.data();
      D.isb = isb.data(), D.sprm = sprm.data(), D.iury = iury.data(), D.uprm = uprm.data(), D.iopb = iopb.data(), D.oprm = oprm.data();
      D.itors = itors.data(), D.tprm = tprm.data(), D.ipit = ipit.data(), D.pprm = pprm.data();
      D.itt = itt.data(), D.ttchk = ttchk.data(), D.ttgrid = ttgrid.data(), D.grids = grids.data();
      D.ttx = ttx.data(), D.tty = tty.data(), D.tbf = tbf.data(), D.tbx = tbx.data(), D.tby = tby.data(), D.tbxy = tbxy.data();
      D.K = K, D.opbtyp = opbtyp;
      return D;
   }
};

// oracle/_ref/libref_valence.so -- the REFERENCE'S OWN valence arithmetic, executed on the CPU.  TEST INFRASTRUCTURE ONLY.
//
// This driver (our code) includes the reference's per-interaction headers where they lie under /root/reference
// (include/seq/{bond,angle,strbnd,urey,opbend,torsion,pitors,tortor}.h, compiled -DTINKER_DOUBLE_PRECISION by
// oracle/Makefile; nothing of the reference is copied into this repository) and calls dk_bond<calc::V1> ... dk_tortor
// <calc::V1> in plain loops, exactly as src/acc/ebond.cpp ... etortor.cpp do inside their OpenACC loops.  It rebuilds the
// index indirections those functions expect (isb -> iang / bl, iopb -> iang, itt -> ibitor) from the resolved atom lists of
// apx_valence.  The oracle (oracle/valence_ref.py) and the CUDA kernel are both held against it at dhfr2 size.
#include "apx.h"
#include "seq/angle.h"
#include "seq/bond.h"
#include "seq/opbend.h"
#include "seq/pitors.h"
#include "seq/strbnd.h"
#include "seq/torsion.h"
#include "seq/tortor.h"
#include "seq/urey.h"
#include <cmath>
#include <cstring>
#include <vector>

using namespace tinker;

namespace {
struct Sums {
   double* e8;
   double vir[6];
   void add(int t, real e, real vxx, real vyx, real vzx, real vyy, real vzy, real vzz)
   {
      e8[t] += e;
      vir[0] += vxx, vir[1] += vyx, vir[2] += vzx, vir[3] += vyy, vir[4] += vzy, vir[5] += vzz;
   }
};
}

extern "C" int ref_valence_eval(const apx_valence* v, const double* xyz, double* e8, double* grad, double* vir9)
{
   typedef calc::V1 Ver;
   const int n = v->n;
   std::vector<real> x(n), y(n), z(n), gx(n, 0), gy(n, 0), gz(n, 0);
   for (int i = 0; i < n; ++i)
      x[i] = xyz[3 * i], y[i] = xyz[3 * i + 1], z[i] = xyz[3 * i + 2];
   std::memset(e8, 0, sizeof(double) * 8);
   Sums S;
   S.e8 = e8;
   std::memset(S.vir, 0, sizeof(S.vir));
   const double* K = v->consts;
   real e, vxx, vyx, vzx, vyy, vzy, vzz;

   if (v->use[0] && v->nbond) {      // src/acc/ebond.cpp
      std::vector<real> bl(v->bl, v->bl + v->nbond), bk(v->bk, v->bk + v->nbond);
      auto ibnd = reinterpret_cast<const int(*)[2]>(v->ibnd);
      for (int i = 0; i < v->nbond; ++i) {
         dk_bond<Ver>(e, vxx, vyx, vzx, vyy, vzy, vzz, gx.data(), gy.data(), gz.data(), Bond::HARMONIC, (real)K[0], i, ibnd, bl.data(), bk.data(),
            (real)K[1], (real)K[2], x.data(), y.data(), z.data());
         S.add(0, e, vxx, vyx, vzx, vyy, vzy, vzz);
      }
   }
   if (v->use[1] && v->nangle) {      // src/acc/eangle.cpp
      std::vector<Angle> typ(v->nangle);
      std::vector<real> anat(v->anat, v->anat + v->nangle), ak(v->ak, v->ak + v->nangle), afld(v->nangle, 0);
      for (int i = 0; i < v->nangle; ++i)
         typ[i] = v->angtyp[i] == 1 ? Angle::IN_PLANE : Angle::HARMONIC;
      auto iang = reinterpret_cast<const int(*)[4]>(v->iang);
      for (int i = 0; i < v->nangle; ++i) {
         dk_angle<Ver>(e, vxx, vyx, vzx, vyy, vzy, vzz, gx.data(), gy.data(), gz.data(), typ.data(), (real)K[3], i, iang, anat.data(), ak.data(),
            afld.data(), (real)K[4], (real)K[5], (real)K[6], (real)K[7], x.data(), y.data(), z.data());
         S.add(1, e, vxx, vyx, vzx, vyy, vzy, vzz);
      }
   }
   if (v->use[2] && v->nstrbnd) {      // src/acc/estrbnd.cpp: isb = {angle, bond a-b, bond c-b}
      const int m = v->nstrbnd;
      std::vector<int> iang(4 * m), isb(3 * m);
      std::vector<real> anat(m), bl(2 * m), sbk(2 * m);
      for (int i = 0; i < m; ++i) {
         iang[4 * i] = v->isb[3 * i], iang[4 * i + 1] = v->isb[3 * i + 1], iang[4 * i + 2] = v->isb[3 * i + 2], iang[4 * i + 3] = 0;
         isb[3 * i] = i, isb[3 * i + 1] = 2 * i, isb[3 * i + 2] = 2 * i + 1;
         anat[i] = v->sb_anat[i], bl[2 * i] = v->sb_bl[2 * i], bl[2 * i + 1] = v->sb_bl[2 * i + 1];
         sbk[2 * i] = v->sbk[2 * i], sbk[2 * i + 1] = v->sbk[2 * i + 1];
      }
      for (int i = 0; i < m; ++i) {
         dk_strbnd<Ver>(e, vxx, vyx, vzx, vyy, vzy, vzz, gx.data(), gy.data(), gz.data(), (real)K[8], i, reinterpret_cast<const int(*)[3]>(isb.data()),
            reinterpret_cast<const real(*)[2]>(sbk.data()), bl.data(), reinterpret_cast<const int(*)[4]>(iang.data()), anat.data(), x.data(), y.data(),
            z.data());
         S.add(2, e, vxx, vyx, vzx, vyy, vzy, vzz);
      }
   }
   if (v->use[3] && v->nurey) {      // src/acc/eurey.cpp
      std::vector<real> uk(v->uk, v->uk + v->nurey), ul(v->ul, v->ul + v->nurey);
      auto iury = reinterpret_cast<const int(*)[3]>(v->iury);
      for (int i = 0; i < v->nurey; ++i) {
         dk_urey<Ver>(e, vxx, vyx, vzx, vyy, vzy, vzz, gx.data(), gy.data(), gz.data(), (real)K[9], i, iury, uk.data(), ul.data(), (real)K[10],
            (real)K[11], x.data(), y.data(), z.data());
         S.add(3, e, vxx, vyx, vzx, vyy, vzy, vzz);
      }
   }
   if (v->use[4] && v->nopbend) {      // src/acc/eopbend.cpp: iopb = angle index, iang[.][3] = out-of-plane atom
      const int m = v->nopbend;
      std::vector<int> iopb(m);
      std::vector<real> opbk(v->opbk, v->opbk + m);
      for (int i = 0; i < m; ++i)
         iopb[i] = i;
      auto iang = reinterpret_cast<const int(*)[4]>(v->iopb);
      for (int i = 0; i < m; ++i) {
         dk_opbend<Ver>(e, vxx, vyx, vzx, vyy, vzy, vzz, gx.data(), gy.data(), gz.data(), v->opbtyp == 1 ? OPBend::ALLINGER : OPBend::WDC, (real)K[12],
            i, iopb.data(), opbk.data(), iang, (real)K[13], (real)K[14], (real)K[15], (real)K[16], x.data(), y.data(), z.data());
         S.add(4, e, vxx, vyx, vzx, vyy, vzy, vzz);
      }
   }
   if (v->use[5] && v->ntors) {      // src/acc/etors.cpp: torsN = {amplitude, phase, cos, sin}
      const int m = v->ntors;
      std::vector<real> t[6];
      const double deg = 3.141592653589793238 / 180.0;
      for (int k = 0; k < 6; ++k) {
         t[k].resize(4 * (size_t)m);
         for (int i = 0; i < m; ++i) {
            const double ph = v->tors_phase[6 * i + k];
            t[k][4 * i] = v->tors_v[6 * i + k], t[k][4 * i + 1] = ph, t[k][4 * i + 2] = std::cos(ph * deg), t[k][4 * i + 3] = std::sin(ph * deg);
         }
      }
      auto itors = reinterpret_cast<const int(*)[4]>(v->itors);
      auto T = [&](int k) { return reinterpret_cast<const real(*)[4]>(t[k].data()); };
      for (int i = 0; i < m; ++i) {
         dk_tors<Ver>(e, vxx, vyx, vzx, vyy, vzy, vzz, gx.data(), gy.data(), gz.data(), (real)K[17], i, itors, T(0), T(1), T(2), T(3), T(4), T(5),
            x.data(), y.data(), z.data());
         S.add(5, e, vxx, vyx, vzx, vyy, vzy, vzz);
      }
   }
   if (v->use[6] && v->npitors) {      // src/acc/epitors.cpp
      std::vector<real> kpit(v->kpit, v->kpit + v->npitors);
      auto ipit = reinterpret_cast<const int(*)[6]>(v->ipit);
      for (int i = 0; i < v->npitors; ++i) {
         dk_pitors<Ver>(e, vxx, vyx, vzx, vyy, vzy, vzz, gx.data(), gy.data(), gz.data(), (real)K[18], i, ipit, kpit.data(), x.data(), y.data(), z.data());
         S.add(6, e, vxx, vyx, vzx, vyy, vzy, vzz);
      }
   }
   if (v->use[7] && v->ntortor) {      // src/acc/etortor.cpp: itt = {bitorsion, grid, flag}; the grids as ktrtor.f dimensions them
      const int m = v->ntortor, ng = v->ngrid;
      const int MG = ktrtor::maxtgrd, MG2 = ktrtor::maxtgrd2;
      std::vector<int> itt(3 * m), tnx(v->tnx, v->tnx + ng), tny(v->tny, v->tny + ng), chk(v->tt_chk, v->tt_chk + m);
      std::vector<real> ttx((size_t)ng * MG, 0), tty((size_t)ng * MG, 0), tbf((size_t)ng * MG2, 0), tbx((size_t)ng * MG2, 0), tby((size_t)ng * MG2, 0),
         tbxy((size_t)ng * MG2, 0);
      for (int g = 0; g < ng; ++g) {
         if (tnx[g] > MG || tny[g] > MG)
            return 2;
         for (int k = 0; k < tnx[g]; ++k)
            ttx[(size_t)g * MG + k] = v->ttx[v->tt_xoff[g] + k];
         for (int k = 0; k < tny[g]; ++k)
            tty[(size_t)g * MG + k] = v->tty[v->tt_yoff[g] + k];
         for (int k = 0; k < tnx[g] * tny[g]; ++k) {
            tbf[(size_t)g * MG2 + k] = v->tbf[v->tt_off[g] + k], tbx[(size_t)g * MG2 + k] = v->tbx[v->tt_off[g] + k];
            tby[(size_t)g * MG2 + k] = v->tby[v->tt_off[g] + k], tbxy[(size_t)g * MG2 + k] = v->tbxy[v->tt_off[g] + k];
         }
      }
      for (int i = 0; i < m; ++i)
         itt[3 * i] = i, itt[3 * i + 1] = v->tt_grid[i], itt[3 * i + 2] = 0;      // 0: atoms already in table order
      auto ibitor = reinterpret_cast<const int(*)[5]>(v->itt);
      for (int i = 0; i < m; ++i) {
         dk_tortor<Ver>(e, vxx, vyx, vzx, vyy, vzy, vzz, gx.data(), gy.data(), gz.data(), (real)K[19], i, reinterpret_cast<const int(*)[3]>(itt.data()),
            ibitor, chk.data(), tnx.data(), tny.data(), reinterpret_cast<const real(*)[ktrtor::maxtgrd]>(ttx.data()),
            reinterpret_cast<const real(*)[ktrtor::maxtgrd]>(tty.data()), reinterpret_cast<const real(*)[ktrtor::maxtgrd2]>(tbf.data()),
            reinterpret_cast<const real(*)[ktrtor::maxtgrd2]>(tbx.data()), reinterpret_cast<const real(*)[ktrtor::maxtgrd2]>(tby.data()),
            reinterpret_cast<const real(*)[ktrtor::maxtgrd2]>(tbxy.data()), x.data(), y.data(), z.data());
         S.add(7, e, vxx, vyx, vzx, vyy, vzy, vzz);
      }
   }
   for (int i = 0; i < n; ++i)
      grad[3 * i] = gx[i], grad[3 * i + 1] = gy[i], grad[3 * i + 2] = gz[i];
   const double* s = S.vir;
   const double m9[9] = {s[0], s[1], s[2], s[1], s[3], s[4], s[2], s[4], s[5]};
   std::memcpy(vir9, m9, sizeof(m9));
   return 0;
}

// CPU harness for the valence math of the CUDA library -- TEST INFRASTRUCTURE.  Compiles the very headers the kernel
// uses (csrc/valmath.cuh, valterms.cuh, valpack.h) with g++ and walks the interactions in a plain loop, so the
// hand-derived gradients can be held against the oracle without a GPU (tests/test_valence_math.py).
#include "valpack.h"
#include <cstring>

namespace {
struct HostAcc {
   double* e8;
   double* g;
   double* vir6;
   void energy(int term, double e) { e8[term] += e; }
   template <class R>
   void grad(int atom, R x, R y, R z) { g[3 * atom] += (double)x, g[3 * atom + 1] += (double)y, g[3 * atom + 2] += (double)z; }
   template <class R>
   void virial(const R* v)
   {
      for (int k = 0; k < 6; ++k)
         vir6[k] += (double)v[k];
   }
};

template <class R>
void run(const apx_valence* v, const double* xyz, int do_v, double* e8, double* grad, double* vir6)
{
   ValPacked<R> P(*v);
   vm::ValDev<R> D = P.view();
   HostAcc acc{e8, grad, vir6};
   for (int idx = 0; idx < D.off[vm::T_COUNT]; ++idx)
      vm::eval_interaction<R>(D, idx, xyz, true, do_v != 0, acc);
}
}

extern "C" int valmath_host_eval(int real_bytes, const apx_valence* v, const double* xyz, int do_v, double* e8, double* grad, double* vir6)
{
   std::memset(e8, 0, sizeof(double) * 8);
   std::memset(grad, 0, sizeof(double) * 3 * (size_t)v->n);
   std::memset(vir6, 0, sizeof(double) * 6);
   if (real_bytes == 4)
      run<float>(v, xyz, do_v, e8, grad, vir6);
   else
      run<double>(v, xyz, do_v, e8, grad, vir6);
   return 0;
}

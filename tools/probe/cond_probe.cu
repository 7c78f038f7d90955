// Probe: does this driver run a conditional WHILE graph node whose body (captured from two streams with an event fork/join, like
// one PCG iteration) ends the loop from the device with cudaGraphSetConditional, and does the default value come back on
// every launch?  Prints the loop count of three launches (expect 5 5 5) and the time per launch.
#include <cuda_runtime.h>
#include <cstdio>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)
__global__ void k_a(int* counter) { if (threadIdx.x == 0 && blockIdx.x == 0) atomicAdd(counter, 1); }
__global__ void k_side(float* x) { x[threadIdx.x] += 1.0f; }
__global__ void k_end(int* counter, cudaGraphConditionalHandle h)
{
   if (*counter >= 5)
      cudaGraphSetConditional(h, 0);
}
int main()
{
   cudaStream_t st, s2;
   CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
   CK(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
   cudaEvent_t fork, join, t0, t1;
   CK(cudaEventCreateWithFlags(&fork, cudaEventDisableTiming));
   CK(cudaEventCreateWithFlags(&join, cudaEventDisableTiming));
   CK(cudaEventCreate(&t0));
   CK(cudaEventCreate(&t1));
   int* d;
   float* x;
   CK(cudaMalloc(&d, 4));
   CK(cudaMalloc(&x, 128));
   CK(cudaMemset(x, 0, 128));
   cudaGraph_t g;
   CK(cudaGraphCreate(&g, 0));
   cudaGraphConditionalHandle h;
   CK(cudaGraphConditionalHandleCreate(&h, g, 1, cudaGraphCondAssignDefault));
   cudaGraphNodeParams p = {cudaGraphNodeTypeConditional};
   p.conditional.handle = h;
   p.conditional.type = cudaGraphCondTypeWhile;
   p.conditional.size = 1;
   cudaGraphNode_t node;
   CK(cudaGraphAddNode(&node, g, nullptr, 0, &p));
   cudaGraph_t body = p.conditional.phGraph_out[0];
   CK(cudaStreamBeginCaptureToGraph(st, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
   k_a<<<4, 32, 0, st>>>(d);
   CK(cudaEventRecord(fork, st));
   CK(cudaStreamWaitEvent(s2, fork, 0));
   k_side<<<1, 32, 0, s2>>>(x);
   CK(cudaEventRecord(join, s2));
   CK(cudaMemsetAsync(x + 16, 0, 16, st));
   CK(cudaStreamWaitEvent(st, join, 0));
   k_end<<<1, 1, 0, st>>>(d, h);
   cudaGraph_t out;
   CK(cudaStreamEndCapture(st, &out));
   cudaGraphExec_t ex;
   CK(cudaGraphInstantiate(&ex, g, 0));
   for (int rep = 0; rep < 3; ++rep) {
      CK(cudaMemsetAsync(d, 0, 4, st));
      CK(cudaEventRecord(t0, st));
      CK(cudaGraphLaunch(ex, st));
      CK(cudaEventRecord(t1, st));
      CK(cudaStreamSynchronize(st));
      int hv = -1;
      float ms = 0;
      CK(cudaMemcpy(&hv, d, 4, cudaMemcpyDeviceToHost));
      cudaEventElapsedTime(&ms, t0, t1);
      printf("launch %d: loop ran %d times (expect 5), %.1f us per iteration\n", rep, hv, 1000 * ms / (hv > 0 ? hv : 1));
   }
   float hx = 0;
   CK(cudaMemcpy(&hx, x, 4, cudaMemcpyDeviceToHost));
   printf("side-stream kernel ran %d times (expect 15)\n", (int)hx);
   printf("cond_probe OK\n");
   return 0;
}

// graphwhile.cu -- feasibility probe for the graph-driven step loop: a CUDA graph with an outer WHILE
// node (routing steps) whose body holds an inner WHILE node (Picard trials), conditions set from
// device code.  Prints the per-iteration overhead of the empty loop structure.
//   nvcc -arch=sm_100a -o graphwhile graphwhile.cu && ./graphwhile
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s -> %s (line %d)\n", #x, cudaGetErrorString(e), __LINE__); return 1; } } while (0)

struct Ctl { int step, nSteps, trial, maxTrials; unsigned long long work; };

__global__ void step_begin(Ctl *c, cudaGraphConditionalHandle inner)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) { c->trial = 0; cudaGraphSetConditional(inner, 1); }
}
__global__ void trial_body(Ctl *c)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) atomicAdd(&c->work, 1ull);
}
__global__ void trial_ctl(Ctl *c, cudaGraphConditionalHandle inner)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        c->trial++;
        // data-dependent trip count: 2..4 trials depending on the step
        int want = 2 + (c->step % 3);
        cudaGraphSetConditional(inner, c->trial < want && c->trial < c->maxTrials);
    }
}
__global__ void step_end(Ctl *c, cudaGraphConditionalHandle outer)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        c->step++;
        cudaGraphSetConditional(outer, c->step < c->nSteps);
    }
}

int main()
{
    Ctl *c;
    CK(cudaMalloc(&c, sizeof(Ctl)));
    cudaStream_t s;
    CK(cudaStreamCreate(&s));
    cudaGraph_t g;
    CK(cudaGraphCreate(&g, 0));
    cudaGraphConditionalHandle hOuter, hInner;
    CK(cudaGraphConditionalHandleCreate(&hOuter, g, 1, cudaGraphCondAssignDefault));
    CK(cudaGraphConditionalHandleCreate(&hInner, g, 0, 0));

    cudaGraphNodeParams po = {};
    po.type = cudaGraphNodeTypeConditional;
    po.conditional.handle = hOuter;
    po.conditional.type = cudaGraphCondTypeWhile;
    po.conditional.size = 1;
    cudaGraphNode_t nOuter;
    CK(cudaGraphAddNode(&nOuter, g, nullptr, 0, &po));
    cudaGraph_t body = po.conditional.phGraph_out[0];

    // body: step_begin -> inner while {trial_body x2 -> trial_ctl} -> step_end
    auto add_kernel = [&](cudaGraph_t gr, cudaGraphNode_t *dep, int ndep, void *fn, void **args, int blocks,
                          cudaGraphNode_t *out) {
        cudaKernelNodeParams kp = {};
        kp.func = fn; kp.gridDim = dim3(blocks); kp.blockDim = dim3(128); kp.kernelParams = args;
        return cudaGraphAddKernelNode(out, gr, dep, ndep, &kp);
    };
    cudaGraphNode_t nBegin, nInner, nEnd, nB1, nB2, nCtl;
    void *aBegin[] = { &c, &hInner };
    CK(add_kernel(body, nullptr, 0, (void *)step_begin, aBegin, 1, &nBegin));
    cudaGraphNodeParams pi = {};
    pi.type = cudaGraphNodeTypeConditional;
    pi.conditional.handle = hInner;
    pi.conditional.type = cudaGraphCondTypeWhile;
    pi.conditional.size = 1;
    CK(cudaGraphAddNode(&nInner, body, &nBegin, 1, &pi));
    cudaGraph_t inner = pi.conditional.phGraph_out[0];
    void *aBody[] = { &c };
    CK(add_kernel(inner, nullptr, 0, (void *)trial_body, aBody, 296, &nB1));
    CK(add_kernel(inner, &nB1, 1, (void *)trial_body, aBody, 296, &nB2));
    void *aCtl[] = { &c, &hInner };
    CK(add_kernel(inner, &nB2, 1, (void *)trial_ctl, aCtl, 1, &nCtl));
    void *aEnd[] = { &c, &hOuter };
    CK(add_kernel(body, &nInner, 1, (void *)step_end, aEnd, 1, &nEnd));

    cudaGraphExec_t ex;
    CK(cudaGraphInstantiate(&ex, g, 0));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 3; rep++) {
        Ctl h = { 0, 300, 0, 8, 0ull };
        CK(cudaMemcpy(c, &h, sizeof(h), cudaMemcpyHostToDevice));
        CK(cudaEventRecord(e0, s));
        CK(cudaGraphLaunch(ex, s));
        CK(cudaEventRecord(e1, s));
        CK(cudaStreamSynchronize(s));
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        CK(cudaMemcpy(&h, c, sizeof(h), cudaMemcpyDeviceToHost));
        // expected trials: sum over steps of 2 + step % 3 = 300 * 3 = 900; work = 2 * 900
        printf("rep %d: steps %d work %llu (expect 1800) %.3f ms -> %.2f us per trial (3 kernels), %.2f us per kernel\n",
               rep, h.step, h.work, ms, 1000.0 * ms / 900, 1000.0 * ms / (900 * 3 + 600));
    }
    return 0;
}

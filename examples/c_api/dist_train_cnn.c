/*
 * Data-parallel training of the demo CNN over the HiPS parameter server, every process a plain C program — the flow of the reference's
 * examples/cnn.py (:96-131: init keys, per step forward / backward, push(idx, grad, priority=-idx) + pull(idx) for every parameter) without
 * Python:  workers compute with GXExecutor*, the server aggregates the workers' gradients and applies the optimizer natively (the spec
 * travels as server command 7, what kv.set_optimizer sends from Python), the scheduler does rendezvous and barriers.
 *
 *   gcc -O2 -I geomx_b200/include examples/c_api/dist_train_cnn.c -L geomx_b200/lib -lgeomx_capi -Wl,-rpath,$PWD/geomx_b200/lib -lm -o dist_train_cnn
 *   export DMLC_PS_ROOT_URI=127.0.0.1 DMLC_PS_ROOT_PORT=9092 DMLC_NUM_SERVER=1 DMLC_NUM_WORKER=2 DMLC_NUM_ALL_WORKER=2
 *   DMLC_ROLE=scheduler ./dist_train_cnn & DMLC_ROLE=server ./dist_train_cnn & DMLC_ROLE=worker ./dist_train_cnn 40 & DMLC_ROLE=worker ./dist_train_cnn 40
 */
#define BATCH 16
#include "cnn_common.h"

#define KCK(call) do { if ((call) != 0) { fprintf(stderr, "%s:%d: %s\n", __FILE__, __LINE__, GXGetLastError()); exit(1); } } while (0)

int main(int argc, char** argv) {
  const int steps = argc > 1 ? atoi(argv[1]) : 40;
  KVStoreHandle kv;
  int is_worker = 0, rank = 0, nworkers = 1;
  KCK(GXKVStoreIsWorkerNode(&is_worker));
  KCK(GXKVStoreCreate("dist_sync", &kv));
  if (!is_worker) {                                        /* scheduler / server: serve until the workers end the job */
    KCK(GXKVStoreRunServer(kv));
    KCK(GXKVStoreFree(kv));
    return 0;
  }
  KCK(GXKVStoreGetRank(kv, &rank)); KCK(GXKVStoreGetGroupSize(kv, &nworkers));
  if (rank == 0) KCK(GXKVStoreSendCommmandToServers(kv, 7, "name=sgd;lr=0.1;wd=0.0;rescale_grad=1.0;clip_gradient=-1.0;momentum=0.0"));
  {
    SymbolHandle net = build();
    uint32_t nargs, naux, i, nout;
    const char** names;
    const char* shape_keys[] = {"data"};
    const uint32_t ind[] = {0, 4}, dims[] = {B, 1, 28, 28};
    const char* no_grad[] = {"data", "softmax_label"};
    ExecutorHandle ex;
    NDArrayHandle *args, *grads, *aux, *outs;
    NDArrayHandle arg_copy[16], grad_copy[16];
    int is_param[16], data_i = -1, label_i = -1, step, handle;
    static float X[B * 784], y[B], prob[B * 10];
    float first = 0, last = 0;
    double checksum = 0;

    CK(GXExecutorSimpleBind(net, 1, shape_keys, ind, dims, "write", 2, no_grad, &ex, &nargs, &args, &grads, &naux, &aux));
    memcpy(arg_copy, args, nargs * sizeof(NDArrayHandle)); memcpy(grad_copy, grads, nargs * sizeof(NDArrayHandle));
    CK(GXSymbolListArguments(net, &nargs, &names));
    for (i = 0; i < nargs; ++i) {
      is_param[i] = 1;
      if (!strcmp(names[i], "data")) { data_i = (int)i; is_param[i] = 0; }
      if (!strcmp(names[i], "softmax_label")) { label_i = (int)i; is_param[i] = 0; }
    }
    /* every worker draws its own initial values; Init keeps rank 0's, the pull below makes everybody start from them */
    rng_state = 777u + 1000u * (uint32_t)rank;
    for (i = 0; i < nargs; ++i) {
      float* w; size_t n, k; uint32_t nd; const uint32_t* s;
      if (!is_param[i]) continue;
      n = numel(arg_copy[i]);
      CK(GXNDArrayGetData(arg_copy[i], (void**)&w)); CK(GXNDArrayGetShape(arg_copy[i], &nd, &s));
      for (k = 0; k < n; ++k) w[k] = nd > 1 ? (2.0f * frand() - 1.0f) * sqrtf(3.0f / (float)(n / s[0])) : 0.0f;
      KCK(GXKVStoreInit(kv, (int)i, w, n, 0));
      KCK(GXKVStorePull(kv, (int)i, w, n, 0, 0, &handle)); KCK(GXKVStoreWait(kv, handle));
    }
    rng_state = 4242u + 99u * (uint32_t)rank;                /* different data on every worker */
    for (step = 0; step < steps; ++step) {
      float loss = 0; int b, pulls[16];
      make_batch(X, y);
      CK(GXNDArraySyncCopyFromCPU(arg_copy[data_i], X, B * 784)); CK(GXNDArraySyncCopyFromCPU(arg_copy[label_i], y, B));
      CK(GXExecutorForward(ex, 1)); CK(GXExecutorBackward(ex, 0, NULL));
      CK(GXExecutorOutputs(ex, &nout, &outs)); CK(GXNDArraySyncCopyToCPU(outs[0], prob, B * 10));
      for (b = 0; b < B; ++b) loss -= logf(prob[b * 10 + (int)y[b]] + 1e-12f) / B;
      if (step == 0) first = loss;
      last = loss;
      for (i = nargs; i-- > 0;) {                             /* last layer first, like priority = -idx: its gradient is ready first */
        float *w, *g; size_t n, k;
        if (!is_param[i]) continue;
        n = numel(arg_copy[i]);
        CK(GXNDArrayGetData(arg_copy[i], (void**)&w)); CK(GXNDArrayGetData(grad_copy[i], (void**)&g));
        for (k = 0; k < n; ++k) g[k] /= (float)nworkers;      /* the server sums the workers' pushes */
        KCK(GXKVStorePush(kv, (int)i, g, n, 0, -(int)i, &handle));
        KCK(GXKVStorePull(kv, (int)i, w, n, 0, -(int)i, &pulls[i]));
      }
      for (i = 0; i < nargs; ++i) if (is_param[i]) KCK(GXKVStoreWait(kv, pulls[i]));
      if (step % 10 == 0) printf("rank %d step %d loss %.4f\n", rank, step, loss);
    }
    for (i = 0; i < nargs; ++i) {
      float* w; size_t n, k;
      if (!is_param[i]) continue;
      n = numel(arg_copy[i]);
      CK(GXNDArrayGetData(arg_copy[i], (void**)&w));
      for (k = 0; k < n; ++k) checksum += (double)w[k] * (double)(1 + (k + i) % 7);
    }
    printf("FINAL rank %d of %d loss %.4f -> %.4f checksum %.6f\n", rank, nworkers, first, last, checksum);
    CK(GXExecutorFree(ex)); CK(GXSymbolFree(net));
  }
  KCK(GXKVStoreFree(kv));
  return 0;
}

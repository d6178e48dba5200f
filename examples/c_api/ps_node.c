/*
 * One HiPS process (scheduler, server or worker — chosen by DMLC_ROLE like every node of the reference) written against the flat C ABI only.
 * The server installs its optimizer as a C callback (role of MXKVStoreSetUpdater on a server process, kvstore_dist_server.h:346-375):
 *     weight -= 0.1 * aggregated_gradient
 * Workers init key 3, push rank-dependent gradients twice and pull; each prints the pulled value per step.
 *
 *   gcc -O2 -I geomx_b200/include examples/c_api/ps_node.c -L geomx_b200/lib -lgeomx_capi -Wl,-rpath,$PWD/geomx_b200/lib -o ps_node
 *   DMLC_ROLE=scheduler DMLC_PS_ROOT_URI=127.0.0.1 DMLC_PS_ROOT_PORT=9091 DMLC_NUM_SERVER=1 DMLC_NUM_WORKER=2 ./ps_node &   (same for server, worker x2)
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <geomx/c_api.h>

#define CK(call) do { if ((call) != 0) { fprintf(stderr, "%s:%d: %s\n", __FILE__, __LINE__, GXGetLastError()); exit(1); } } while (0)
#define N 6

static int updates = 0, commands = 0;

static void sgd(int key, const float* grad, float* weight, size_t n, void* arg) {
  const float lr = *(const float*)arg; size_t i;
  (void)key;
  for (i = 0; i < n; ++i) weight[i] -= lr * grad[i];
  ++updates;
}
static void controller(int head, const char* body, void* arg) { (void)arg; (void)body; if (head == 0) ++commands; }

int main(void) {
  const char* role = getenv("DMLC_ROLE");
  KVStoreHandle kv;
  int is_worker = 0;
  CK(GXKVStoreIsWorkerNode(&is_worker));
  CK(GXKVStoreCreate("dist_sync", &kv));
  if (!is_worker) {
    float lr = 0.1f;
    CK(GXKVStoreRunServerEx(kv, controller, NULL, sgd, &lr));          /* blocks until the workers shut the job down */
    printf("%s done: %d updates, %d controller commands\n", role ? role : "?", updates, commands);
    CK(GXKVStoreFree(kv));
    return 0;
  }
  {
    int rank, nw, step, hp, hl, i;
    const char* type;
    float w[N], g[N];
    CK(GXKVStoreGetRank(kv, &rank)); CK(GXKVStoreGetGroupSize(kv, &nw)); CK(GXKVStoreGetType(kv, &type));
    if (rank == 0) CK(GXKVStoreSendCommmandToServers(kv, 0, "hello from C"));
    for (i = 0; i < N; ++i) w[i] = 1.0f;
    CK(GXKVStoreInit(kv, 3, w, N, 0));
    for (step = 0; step < 2; ++step) {
      for (i = 0; i < N; ++i) g[i] = 0.5f * (float)(rank + 1);
      CK(GXKVStorePush(kv, 3, g, N, 0, 0, &hp));
      CK(GXKVStorePull(kv, 3, w, N, 0, 0, &hl));
      CK(GXKVStoreWait(kv, hl));
      printf("RESULT rank %d of %d type %s step %d value %.6f\n", rank, nw, type, step, w[0]);
    }
    CK(GXKVStoreFree(kv));
  }
  return 0;
}

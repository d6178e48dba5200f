/*
 * Shared pieces of the C examples: the reference's demo CNN (examples/cnn.py:56-66) built through GXSymbol*, a self-contained synthetic
 * data set (one bright 7x7 patch per class on noise), a small LCG.  Include after <geomx/c_api.h>; define BATCH before including.
 */
#ifndef GEOMX_EXAMPLES_CNN_COMMON_H_
#define GEOMX_EXAMPLES_CNN_COMMON_H_
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <geomx/c_api.h>

#ifndef BATCH
#define BATCH 32
#endif
#define B BATCH

#define CK(call) do { if ((call) != 0) { fprintf(stderr, "%s:%d: %s\n", __FILE__, __LINE__, GXRTGetLastError()); exit(1); } } while (0)

static uint32_t rng_state = 12345u;
static float frand(void) { rng_state = rng_state * 1664525u + 1013904223u; return (float)(rng_state >> 8) / 16777216.0f; }

static SymbolHandle layer(const char* op, const char* name, SymbolHandle in, int nattr, const char** keys, const char** vals) {
  SymbolHandle s;
  CK(GXSymbolCreateAtomicSymbolByName(op, (uint32_t)nattr, keys, vals, &s));
  CK(GXSymbolCompose(s, name, 1, NULL, &in));
  CK(GXSymbolFree(in));                                   /* the composed node keeps its own reference to the input graph */
  return s;
}

static SymbolHandle build(void) {
  SymbolHandle x, label, h, out;
  CK(GXSymbolCreateVariable("data", &x));
  CK(GXSymbolCreateVariable("softmax_label", &label));
  { const char* k[] = {"kernel", "num_filter"}; const char* v[] = {"(5, 5)", "16"}; h = layer("Convolution", "conv0", x, 2, k, v); }
  { const char* k[] = {"act_type"}; const char* v[] = {"relu"}; h = layer("Activation", "relu0", h, 1, k, v); }
  { const char* k[] = {"kernel", "stride", "pool_type"}; const char* v[] = {"(2, 2)", "(2, 2)", "max"}; h = layer("Pooling", "pool0", h, 3, k, v); }
  { const char* k[] = {"kernel", "num_filter"}; const char* v[] = {"(5, 5)", "32"}; h = layer("Convolution", "conv1", h, 2, k, v); }
  { const char* k[] = {"act_type"}; const char* v[] = {"relu"}; h = layer("Activation", "relu1", h, 1, k, v); }
  { const char* k[] = {"kernel", "stride", "pool_type"}; const char* v[] = {"(2, 2)", "(2, 2)", "max"}; h = layer("Pooling", "pool1", h, 3, k, v); }
  h = layer("Flatten", "flat", h, 0, NULL, NULL);
  { const char* k[] = {"num_hidden"}; const char* v[] = {"256"}; h = layer("FullyConnected", "fc0", h, 1, k, v); }
  { const char* k[] = {"act_type"}; const char* v[] = {"relu"}; h = layer("Activation", "relu2", h, 1, k, v); }
  { const char* k[] = {"num_hidden"}; const char* v[] = {"128"}; h = layer("FullyConnected", "fc1", h, 1, k, v); }
  { const char* k[] = {"act_type"}; const char* v[] = {"relu"}; h = layer("Activation", "relu3", h, 1, k, v); }
  { const char* k[] = {"num_hidden"}; const char* v[] = {"10"}; h = layer("FullyConnected", "fc2", h, 1, k, v); }
  {
    const char* k[] = {"normalization"}; const char* v[] = {"batch"};
    const char* in_keys[] = {"data", "label"}; SymbolHandle ins[2];
    ins[0] = h; ins[1] = label;
    CK(GXSymbolCreateAtomicSymbolByName("SoftmaxOutput", 1, k, v, &out));
    CK(GXSymbolCompose(out, "softmax", 2, in_keys, ins));
    CK(GXSymbolFree(h)); CK(GXSymbolFree(label));
  }
  return out;
}

static void make_batch(float* X, float* y) {
  int b, i, j;
  for (b = 0; b < B; ++b) {
    const int cls = (int)(frand() * 10.0f) % 10, oy = (cls / 4) * 9, ox = (cls % 4) * 7;
    y[b] = (float)cls;
    for (i = 0; i < 784; ++i) X[b * 784 + i] = 0.1f * frand();
    for (i = 0; i < 7; ++i) for (j = 0; j < 7; ++j) X[b * 784 + (oy + i) * 28 + ox + j] += 0.9f;
  }
}

static size_t numel(NDArrayHandle h) {
  uint32_t nd, i; const uint32_t* s; size_t n = 1;
  CK(GXNDArrayGetShape(h, &nd, &s));
  for (i = 0; i < nd; ++i) n *= s[i];
  return n;
}

#endif  /* GEOMX_EXAMPLES_CNN_COMMON_H_ */

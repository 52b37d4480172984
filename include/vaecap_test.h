/* libvaecap -- kernel-level test entries (no reference counterpart).
 *
 * The reference has no operator boundary below `sess.run`, so nothing in yiyang92/vae_captioning binds these.
 * They expose single kernels of the hot path to the parity tests, so that a GEMM / convolution-gradient mismatch is
 * localised to one launch instead of showing up as a loss difference: tests/test_gemm_gpu.py and
 * tests/test_conv_bwd_gpu.py compare them with fp32 matmul / torch conv2d autograd on the same bf16-rounded inputs.
 * All pointers are DEVICE pointers; status codes and vc_last_error() as in vaecap.h. */
#ifndef VAECAP_TEST_H_
#define VAECAP_TEST_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* out[M, N] = act(A x B + bias) on the tcgen05 mainloop. A is [M, K] K-major (a_mn = 0) or [K, M] MN-major (a_mn = 1),
 * B likewise over N; bf16 operands, fp32 accumulation; out fp32 or bf16; atomic = fp32 atomicAdd (split-K).
 * Mirrors every dense contraction of the step: tf.layers.dense / tf.matmul (main.py:94, encoder.py:60-97,
 * decoder.py:111-129) and their gradients. */
int vc_gemm_bf16(const void* A, int a_mn, long long lda, const void* B, int b_mn, long long ldb, void* out, long long ldo,
                 const float* bias, int M, int N, int K, int bn, int splits, int relu, int out_bf16, int atomic,
                 void* stream);

/* CTA-pair policy of the tcgen05 mainloop (cluster of 2, cta_group::2): -1 = automatic (default; VC_PAIR in the
 * environment), 0 = never, 1 = wherever the tile shape allows it. Lets the parity tests run the same GEMM both ways. */
void vc_test_pair_mode(int mode);
/* Dual-N tiles of the same mainloop (two adjacent 256-column n-tiles share one A tile, accumulators fill the 512 TMEM
 * columns): -1 automatic (VC_DUAL in the environment), 0 never, 1 wherever bn = 256 and there are >= 2 n-tiles. */
void vc_test_dual_mode(int mode);

/* Gradients of one 3x3 SAME convolution (tf.nn.conv2d in utils/image_embeddings.py:40-205, differentiated by
 * ops/optimizers.py:49-82): x, dy bf16 NHWC; w fp32 HWIO; dw fp32 [9*Cin, Cout] (zeroed by the call); dx bf16 NHWC. */
int vc_conv3x3_bwd(const void* x, const void* dy, const float* w, float* dw, void* dx, int B, int hw, int cin, int cout,
                   void* stream);

/* Input gradient of one 3x3 SAME convolution with the ReLU derivative of the layer below taken in the GEMM epilogue
 * (the fine-tune backward pass between un-pooled layers, ops/optimizers.py:49-82 through image_embeddings.py:46-205):
 * dx = conv3x3_same(dy, rot180(w)^T) where act > 0, else 0; dbias[cin] += per-channel sums of dx. act: bf16 NHWC
 * [B, hw, hw, cin] (a ReLU output); dbias must be zeroed by the caller. cin a multiple of 64, <= 512. */
int vc_conv3x3_dgrad_relu(const void* dy, const float* w, const void* act, void* dx, float* dbias, int B, int hw, int cin,
                          int cout, void* stream);

/* Derivative of ReLU (+ 2x2/2 max-pool when pooled) with the bias gradient fused: dY = dA routed to the first maximum
 * of each window where the stored post-ReLU activation `out` is positive; db[C] += per-channel sum of dY.
 * (tf.nn.relu / tf.nn.max_pool gradients, image_embeddings.py:46-211.) */
int vc_relu_pool_bwd(const void* dA, const void* out, void* dY, float* db, int B, int hw, int C, int pooled, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VAECAP_TEST_H_ */

// Single translation unit of libffb200.so (keeps the device error word and helper templates in one module).
#include "gemm.cu"
#if defined(FFB_ATT_BN128)
#include "experimental/attention_bn128.cu"   // measured alternative (tools/gpu_variants.sh); not the product kernel
#elif defined(FFB_ATT_SUMMMA)
#include "experimental/attention_summma.cu"  // row sum on the tensor core (tools/gpu_maxfree.sh); not the product kernel, not yet run
#else
#include "attention.cu"
#endif
#if defined(FFB_ATT_SPLIT)
#include "experimental/attention_d128_split.cu"    // column-split softmax on top of the two experiments below; not yet run
#elif defined(FFB_ATT_SUMMMA)
#include "experimental/attention_d128_summma.cu"   // not the product kernel, not yet run
#else
#include "attention_d128.cu"
#endif
#include "elementwise.cu"
#include "final_step.cu"
#include "engine.cu"
#include "flux_engine.cu"
#include "vae_conv.cu"          // VAE decode: first GPU run pending, see the file headers
#include "vae_elementwise.cu"
#include "vae_engine.cu"
#include "wan_elementwise.cu"   // Wan2.1 T2V: first GPU run pending, see the file headers
#include "wan_engine.cu"

// Single translation unit of libffb200.so (keeps the device error word and helper templates in one module).
#include "gemm.cu"
#include "attention.cu"
#include "attention_d128.cu"
#include "elementwise.cu"
#include "final_step.cu"
#include "engine.cu"
#include "flux_engine.cu"

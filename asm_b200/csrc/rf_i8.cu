// rf_i8.cu -- K2 placeholder (replaced by the tcgen05 kernel).
#include "asm_internal.h"

int32_t rf_i8_sums(asm_ctx* ctx, asm_shard) { return asm_fail(ctx, ASM_E_UNSUPPORTED, "ASM_RF_I8 not built yet"); }
int32_t rf_i8_block(asm_ctx* ctx, int32_t, int64_t, int64_t, int32_t, int64_t, int64_t, uint16_t*) {
  return asm_fail(ctx, ASM_E_UNSUPPORTED, "ASM_RF_I8 not built yet");
}

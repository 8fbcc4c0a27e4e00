/*
 * asmsynth.h -- synthetic inputs of the shapes BASELINE.json names.  Test and benchmark
 * infrastructure: libasmsynth.so is NOT part of the product and libasmgpu.so does not contain it.  The library also
 * carries its own copy of the host encoder (asm_encoder_* of include/asmgpu.h), so a generator can hand out
 * dictionary ids without touching the product library.
 */
#ifndef ASMSYNTH_H_
#define ASMSYNTH_H_

#include <stdint.h>

#include "../include/asmgpu.h" /* asm_encoder, ASM_E_* */

#ifdef __cplusplus
extern "C" {
#endif

/* One chain of n_trees rooted binary trees on n_taxa leaves: all chains share the start tree
 * (random-join shape from start_seed); each sample applies `moves_per_sample` Metropolis steps of
 * a random rooted NNI with target exp(-beta * #clades not in the start tree), seeded by
 * chain_seed.  left/right/root as in asm_encoder_add_topologies (any may be NULL);
 * if enc != NULL (an encoder created by THIS library's asm_encoder_create), clade_ids_out
 * [n_trees][n_taxa-1] receives dictionary ids (ids of clades new to the dictionary are handed out in
 * the order the walk creates them). */
int32_t asm_synth_tree_chain(int32_t n_taxa, int64_t n_trees, uint64_t start_seed, uint64_t chain_seed,
                             int32_t moves_per_sample, double beta, int32_t* left, int32_t* right, int32_t* root,
                             asm_encoder* enc, int32_t* clade_ids_out);
/* AR(1) trace x_t = mu + phi (x_{t-1} - mu) + eps_t, eps ~ N(0,1) from a counter-based generator
 * keyed by seed; x_0 drawn from the stationary law. */
int32_t asm_synth_ar1(double* out, int64_t n, double mu, double phi, uint64_t seed);

#ifdef __cplusplus
}
#endif
#endif /* ASMSYNTH_H_ */

// kernels.h -- internal launch API shared by the C-ABI glue (capi.cu) and the engine (engine.cu).
#pragma once

#include "common.cuh"

namespace mli {

// M-tile descriptor used by the encoder / prefill kernels: 64 consecutive positions of one row
struct TileDesc {
    int row;  // batch row
    int j0;   // first position of the tile
};
constexpr int kTileM = 64;

// ---- tile list for new rows (device-side; lengths never leave the GPU) -----------------------
// tiles[0..*n_tiles) covers positions [0, L_r) of every new row r in chunks of kTileM.
// n_new_dev (optional) overrides n_new_host with a device-resident count.
int launch_build_new_row_tiles(mli_ctx* ctx, const int* new_idx, const int* lengths, int n_new_host,
                               const int* n_new_dev, TileDesc* tiles, int* n_tiles, int max_tiles);

// ---- encoder (src/kernels/encoder.cu:102-147 / :56-92) ----------------------------------------
// tokens come from inp[B,S], or (engine) from req_tok[row_req[r]*S + j] when row_req != nullptr;
// tile_m = positions per TileDesc (kTileM, or 16 for the engine's prefill granules)
int launch_paged_encoder_tiles(mli_ctx* ctx, const float* emb, const float* pos, const int* inp,
                               const int* row_req, const int* req_tok, float* const* page_table,
                               const TileDesc* tiles, const int* n_tiles, int max_tiles,
                               const int* lengths, int S, int d, int tile_m = kTileM);
int launch_dense_encoder(mli_ctx* ctx, const float* emb, const float* pos, const int* inp,
                         float* inp_embedding, const int* lengths, const int* new_idx, int S, int d,
                         int n_new);

// ---- exact-order SIMT GEMMs (k-ascending FMA chains == the reference's naive kernels) ----------
int launch_prefill_kv_paged_simt(mli_ctx* ctx, float* const* page_table, const TileDesc* tiles,
                                 const int* n_tiles, int max_tiles, const int* lengths,
                                 const float* wk, const float* wv, int S, int d);
int launch_qkv_latest_paged_simt(mli_ctx* ctx, float* const* page_table, const int* lengths,
                                 const float* wk, const float* wq, const float* wv, float* q_output,
                                 int B, int S, int d);
int launch_logits_simt(mli_ctx* ctx, const float* attn, const float* emb, float* score, int B, int V,
                       int d);
int launch_prefill_kv_dense_simt(mli_ctx* ctx, const float* inp_embedding, const TileDesc* tiles,
                                 const int* n_tiles, int max_tiles, const int* lengths,
                                 const float* wk, const float* wv, float* kt_cache, float* v_cache,
                                 int S, int di, int dn);
int launch_qkv_latest_dense_simt(mli_ctx* ctx, const float* inp_embedding, const int* lengths,
                                 const float* wk, const float* wq, const float* wv, float* kt_cache,
                                 float* v_cache, float* q_output, int B, int S, int di, int dn);

// ---- tcgen05 3xTF32 GEMMs (gemm_tcgen05.cu) ------------------------------------------------------
bool tcgen05_supported(mli_ctx* ctx);
// split/transposed copies of these operands are built once and trusted until unregistered
int tc_register_weights(mli_ctx* ctx, const float* wk, const float* wq, const float* wv, const float* emb,
                        int d, int V);
void tc_unregister_weights(mli_ctx* ctx, const float* wk, const float* emb);
void tc_unregister_all(mli_ctx* ctx);
int launch_prefill_kv_paged_tc(mli_ctx* ctx, float* const* page_table, const TileDesc* tiles,
                               const int* n_tiles, int max_tiles, const int* lengths,
                               const float* wk, const float* wv, int S, int d);
int launch_qkv_latest_paged_tc(mli_ctx* ctx, float* const* page_table, const int* lengths,
                               const float* wk, const float* wq, const float* wv, float* q_output,
                               int B, int S, int d);
// Split-K partial logits: part[z][B][V] for z < *n_split (kMaxLogitSplit * B * V floats); the
// decoder adds the partials in rank order while it scans for the argmax, so the GEMM needs no
// cross-CTA reduction.  act_rows / counts (optional, device): only rows act_rows[0..counts[0]) are
// computed.
constexpr int kMaxLogitSplit = 16;
int launch_logits_tc(mli_ctx* ctx, const float* attn, const float* emb, float* part, int B, int V,
                     int d, int* n_split, const int* act_rows = nullptr, const int* counts = nullptr);
// the engine's merged projection: latest-token K,q,V of the active rows AND prefill K,V of the new
// rows' earlier positions in ONE launch.  act_rows[0..counts[0]) = active rows, gran[0..counts[1]) =
// 16-position granules of the new rows (ignored when use_gran == 0, i.e. forward rounds > 0).
int launch_step_qkv_tc(mli_ctx* ctx, float* const* page_table, const int* lengths, const int* act_rows,
                       const int* counts, const TileDesc* gran, int max_gran, int use_gran,
                       const float* wk, const float* wq, const float* wv, float* q_output, int B, int S,
                       int d, const int* gran_bound = nullptr);

// ---- fused decode attention --------------------------------------------------------------------
int launch_decode_attention_paged(mli_ctx* ctx, const float* q, float* const* page_table,
                                  const int* lengths, float* out, float* softmax_out, int B, int S,
                                  int d);
int launch_decode_attention_dense(mli_ctx* ctx, const float* q, const float* kt_cache,
                                  const float* v_cache, const int* lengths, float* out,
                                  float* softmax_out, int B, int S, int d);
// warp-per-position variant of the single-launch kernel (decode_attention_wp.cu); workspaces as
// prepared by launch_decode_attention_paged
bool attention_wp_supported(int d);
bool attention_wp_usable(mli_ctx* ctx, int B, int d);
int launch_decode_attention_wp(mli_ctx* ctx, const float* q, float* const* page_table, const int* lengths,
                               float* out, float* part_acc, float* part_ml, int* row_done, int B, int S,
                               int d, int min_dyn);
// the reference's three unfused stages as stand-alone launches (API completeness; not on the product path)
int launch_qkt_unfused(mli_ctx* ctx, const float* q, float* const* page_table, const int* lengths, float* qkt,
                       int B, int S, int d);
int launch_softmax_lengths_unfused(mli_ctx* ctx, float* qkt, const int* lengths, int B, int S);
int launch_softmax_v_unfused(mli_ctx* ctx, const float* p, float* const* page_table, float* out,
                             const int* lengths, int B, int S, int d);
double attention_algorithmic_bytes(const int* lengths_host, int B, int d, int kv_bf16 = 0);

// ---- decoder (src/kernels/decoder.cu:25-91, :128-205) -------------------------------------------
// score = n_split partial logit planes [n_split][B][V] (1 = plain logits); score_out (optional)
// receives the summed logits
int launch_paged_decoder(mli_ctx* ctx, const float* score, int n_split, float* score_out,
                         int* decoder_result, int* lengths,
                         float* const* page_table, const float* pos, const float* emb, int B, int V,
                         int S, int d, int n_dec, int i_dec);
int launch_dense_decoder(mli_ctx* ctx, const float* score, int* decoder_result, int* lengths,
                         float* inp_embedding, const float* pos, const float* emb, int B, int V, int S,
                         int d);


// ---- engine internals used by the token gather (comm.cu) ----------------------------------------
int engine_token_table(mli_engine* e, const int** tokens_dev, const int** counts_dev, int* capacity,
                       int* n_sequence, cudaStream_t* engine_stream, mli_ctx** ctx);

}  // namespace mli

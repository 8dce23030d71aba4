/* Packed-weight blob layout shared by the Python packer (fastdiff_b200/weights.py parses this file
 * for the section order) and the CUDA library.
 *
 * blob := header | fp32 data
 *   header: uint64 magic, uint64 version, uint64 n_sections, then n_sections x {uint64 offset, uint64 count}
 *           (offset/count in fp32 elements from the start of the blob; every offset is a multiple of 64
 *           elements = 256 B so any section can be the source of a 16-byte vector load / bulk copy).
 *
 * Section contents (folded weights: w = g*v/||v|| for every weight-normed Conv1d,
 * /root/reference/modules/FastDiff/module/FastDiff_model.py:115-122):
 *   EMB_FREQ  [64]            exp(-j ln(1e4)/63) built in fp32 exactly like util.py:425-427
 *   FC1_WT    [128][512]      fc_t1.weight^T          FC1_B [512]
 *   FC2_WT    [512][512]      fc_t2.weight^T          FC2_B [512]
 *   FIRST_W   [7][32]         first_audio_conv [k][co]   FIRST_B [32]
 *   FINAL_W   [7][32]         final_conv.0 [k][ci]       FINAL_B [1]
 *   DBn_RES_W [32 ci][32 co]  downsample.n.residual_dense   DBn_RES_B [32]
 *   DBn_CONV_W[3][3 k][32 ci][32 co]  downsample.n.conv.{0,1,2}   DBn_CONV_B [3][32]
 *   LBn_FCT_WT[512][80]       lvc_blocks.n.fc_t.weight^T    LBn_FCT_B [80]
 *   LBn_UP_W  [2r k][32 ci][32 co]  lvc_blocks.n.upsample.weight (ConvTranspose1d (ci,co,k))   LBn_UP_B [32]
 *   LBn_CONV_W[4][3 k][32 ci][32 co] lvc_blocks.n.convs.{0..3}    LBn_CONV_B [4][32]
 *   LBn_KPIN_W[5 j][80 ci][64 co]   kernel_predictor.input_conv.0   LBn_KPIN_B [64]
 *   LBn_KPRES_W[6][3 j][64 ci][64 co] kernel_predictor.residual_conv.{1,3,6,8,11,13}   LBn_KPRES_B [6][64]
 *   LBn_KC_W  [192 kk=(j*64+c)][24832 n]   kernel_conv + bias_conv fused along n, permuted so one GEMM row
 *             is, per LVC layer l, the B operand of the location-variable conv as its SWIZZLE_128B K-major smem image:
 *             per tap k a tile of 64 rows (o) x 128 B (32 input channels), the 16-byte chunk c = i/4 of row o stored at
 *             chunk position c ^ (o & 7).  A bulk copy of the 24,576 B lands it in smem ready for tcgen05.mma, and the SIMT
 *             kernel reads one conflict-free LDS.128 per lane.  Followed by the 64 biases:
 *                                n = l*6208 + ((k*64 + o)*8 + ((i/4) ^ (o & 7)))*4 + i%4   <- kernel_conv channel ((l*32+i)*64+o)*3+k
 *                                n = l*6208 + 6144 + o                                     <- bias_conv channel l*64+o
 *             Block 0 (hop 8, SIMT consumer reading straight from HBM) uses PANEL order instead:
 *                                n = l*6208 + ((k*8 + i/4)*64 + o)*4 + i%4   (consecutive lanes o -> consecutive 16-byte vectors)
 *             (channel maps: modules.py:333-342)
 *   LBn_KC_B  [24832]         same permutation of the two bias vectors
 *   LBn_KCT_HI / LBn_KCT_LO [24832 n][192 kk]   the same matrix transposed (K-major rows: the tcgen05 A operand),
 *             split into tf32 pieces  w = hi + lo,  hi = RN_tf32(w), lo = RN_tf32(w - hi)  (low 13 mantissa bits zero)
 *   LBn_UPT_HI / _LO [2r taps][32 co][8 chunks (ci/4) ^ (co&7)][4] (n = 1, 2): lvc_blocks.n.upsample (ConvTranspose1d weight (ci,co,k))
 *             as SWIZZLE_128B K-major tiles Wt[k][co][ci], tf32 pieces (tensor-core upsample)
 *   DB0_CONVT_HI / _LO [3 layers][3 k][32 co][8 chunks (ci/4) ^ (co&7)][4], DB0_REST_HI / _LO [32 co][8 chunks][4]:
 *             downsample.0 convs and its 1x1 residual as SWIZZLE_128B K-major tiles, tf32 pieces (tensor-core DBlock 0)
 *   LBn_CONVT_HI / LBn_CONVT_LO [4 layers][3 k][32 co][8 chunks (ci/4) ^ (co&7)][4]   lvc_blocks.n.convs.* as SWIZZLE_128B K-major tiles, tf32 pieces
 *
 * fp16-piece operands (mode tc_3xf16; two fp16 values per fp32 blob element, little endian).  w*S = hi + lo with
 * hi = RN_f16(w*S), lo = RN_f16(w*S - hi); S = the per-tensor power of two that puts max|w|*S in (8192, 16384] (SCALES16):
 *   LBn_KCT_F16 [2 planes: hi, lo][24832 n][192 kk] fp16   the LBn_KCT matrix (K-major rows of 384 B: the tcgen05 A operand)
 *   LB0_KCT_F16P / LB0_KC_BP  block 0's LB0_KCT_F16 / LB0_KC_B with the rows in the SWIZZLE_128B image order of blocks 1 and 2 instead of
 *             panel order (same values, same scale): the kernel_conv GEMM then writes block 0 as fp16 pieces for the experimental
 *             tensor-core block-0 LVC kernel (option tc_b0)
 *   LBn_CONV_F16 (n = 0, 1, 2; n = 0 is used by the experimental tensor-core block-0 LVC kernel only) [4 layers][3 k][32 co] rows of 128 B = [32 ci hi | 32 ci lo] fp16, the 16-byte chunk c (8 values) of
 *             row co stored at chunk position c ^ (co & 7): SWIZZLE_128B K-major B-operand tiles of the dilated convs
 *   LBn_CONV_F16M (n = 1, 2; k_lvc_p, merged-N form) [4 layers][12 KB], SWIZZLE_128B K-major B-operand tiles:
 *             T01 [64 rows x 128 B]: row R = piece * 32 + co = [tap 0: 32 ci | tap 1: 32 ci] of that piece (hi rows, then lo rows), chunk c at c ^ (R & 7):
 *                 ONE N = 64 MMA per activation piece and K-slice forms the hi- and the lo-weight products of taps 0 and 1;
 *             T2  [32 rows x 128 B]: row co = [32 ci hi | 32 ci lo] of tap 2, chunk c at c ^ (co & 7) (three-pass form)
 *   LB2_UP_F16M [4 phases][64 rows][128 B] (k_upsample_p4): lvc_blocks.2.upsample (ConvTranspose1d weight (ci, co, k), k = 8, stride 4) as merged-N
 *             SWIZZLE_128B tiles per output phase ph (t = 4 m + ph; taps kk1 = (ph + 2) % 4 on input m + (ph + 2) / 4 and kk1 + 4 on the row before):
 *             row R = piece * 32 + co = [tap kk1: 32 ci | tap kk1 + 4: 32 ci] of that piece, chunk c at c ^ (R & 7); scale SCALES16[41]
 *   FIRST_F16U [128 rows = ph * 32 + co][64 B] (k_upsample_p4): first_audio_conv as four Toeplitz tiles over the audio window a_m = audio[4 m - 3 .. 4 m + 6]:
 *             K = 16 values W_ph[i] = first_w[co][i - ph] (0 <= i - ph <= 6), W_ph[10] = first_b[co], else 0; 32 B hi pieces | 32 B lo pieces; scale SCALES16[41]
 *             (SCALES16[41] = the smaller of the two tensors' own scales, so that both products accumulate into one TMEM tile)
 *   LBn_KPW_F16 [28 slots][16 KB]  kernel-predictor hidden stack as B-operand tiles in consumption order (tensor-core k_kp_hidden_tc):
 *             a tile = 64 rows (co) x 128 B, 16-byte chunk c at position c ^ (co & 7).
 *             slots 0..9  : input_conv taps j = 0..4, two slots per tap: {ci 0..63: hi tile 8 KB | lo tile 8 KB},
 *                           {ci 64..79: one tile with rows [16 ci hi (32 B) | 16 ci lo (32 B) | 64 B zero], then 8 KB unused}
 *             slots 10..27: residual convs l = 0..5, taps j = 0..2: {hi tile 8 KB | lo tile 8 KB} over the 64 input channels
 *   SCALES16  [64]  S of: kernel_conv block n at [n]; lvc_blocks.n.convs.l at [4 + 4 n + l];
 *                   kernel predictor of block n: input_conv at [16 + 8 n], residual conv l at [17 + 8 n + l]; FIRST_F16 at [40]; LB2_UP_F16M and FIRST_F16U at [41]
 *   FIRST_F16 [32 co][128 B]  first_audio_conv as the B operand of the skip MMA of LVC block 2 (fd_kernels_lvcp.cuh): per output channel
 *                   K = 16 fp16 [tap 0..6, bias, 0 x 8] hi (chunks 0, 1) | the same lo (chunks 2, 3) | zeros; 16-byte chunk c at c ^ (co & 7)
 */
#ifndef FD_BLOB_H
#define FD_BLOB_H

#define FD_BLOB_MAGIC 0x3142303032444646ULL /* "FFD200B1" */
#define FD_BLOB_VERSION 15ULL

/* The packer reads the names between FD_SECTIONS_BEGIN / FD_SECTIONS_END in this order. */
/* FD_SECTIONS_BEGIN */
#define FD_SECTIONS(X) \
    X(EMB_FREQ) X(FC1_WT) X(FC1_B) X(FC2_WT) X(FC2_B) \
    X(FIRST_W) X(FIRST_B) X(FINAL_W) X(FINAL_B) \
    X(DB0_RES_W) X(DB0_RES_B) X(DB0_CONV_W) X(DB0_CONV_B) \
    X(DB1_RES_W) X(DB1_RES_B) X(DB1_CONV_W) X(DB1_CONV_B) \
    X(DB2_RES_W) X(DB2_RES_B) X(DB2_CONV_W) X(DB2_CONV_B) \
    X(LB0_FCT_WT) X(LB0_FCT_B) X(LB0_UP_W) X(LB0_UP_B) X(LB0_CONV_W) X(LB0_CONV_B) \
    X(LB0_KPIN_W) X(LB0_KPIN_B) X(LB0_KPRES_W) X(LB0_KPRES_B) X(LB0_KC_W) X(LB0_KC_B) X(LB0_KCT_HI) X(LB0_KCT_LO) X(LB0_CONVT_HI) X(LB0_CONVT_LO) \
    X(LB1_FCT_WT) X(LB1_FCT_B) X(LB1_UP_W) X(LB1_UP_B) X(LB1_CONV_W) X(LB1_CONV_B) \
    X(LB1_KPIN_W) X(LB1_KPIN_B) X(LB1_KPRES_W) X(LB1_KPRES_B) X(LB1_KC_W) X(LB1_KC_B) X(LB1_KCT_HI) X(LB1_KCT_LO) X(LB1_CONVT_HI) X(LB1_CONVT_LO) \
    X(LB2_FCT_WT) X(LB2_FCT_B) X(LB2_UP_W) X(LB2_UP_B) X(LB2_CONV_W) X(LB2_CONV_B) \
    X(LB2_KPIN_W) X(LB2_KPIN_B) X(LB2_KPRES_W) X(LB2_KPRES_B) X(LB2_KC_W) X(LB2_KC_B) X(LB2_KCT_HI) X(LB2_KCT_LO) X(LB2_CONVT_HI) X(LB2_CONVT_LO) \
    X(DB0_CONVT_HI) X(DB0_CONVT_LO) X(DB0_REST_HI) X(DB0_REST_LO) \
    X(LB1_UPT_HI) X(LB1_UPT_LO) X(LB2_UPT_HI) X(LB2_UPT_LO) \
    X(LB0_KCT_F16) X(LB1_KCT_F16) X(LB2_KCT_F16) X(LB1_CONV_F16) X(LB2_CONV_F16) \
    X(LB0_KPW_F16) X(LB1_KPW_F16) X(LB2_KPW_F16) X(LB0_CONV_F16) X(LB0_KCT_F16P) X(LB0_KC_BP) X(FIRST_F16) X(LB1_CONV_F16M) X(LB2_CONV_F16M) X(LB2_UP_F16M) X(FIRST_F16U) X(SCALES16)
/* FD_SECTIONS_END */

enum fd_section {
#define FD_X(name) FD_S_##name,
    FD_SECTIONS(FD_X)
#undef FD_X
    FD_S_COUNT
};

/* stride between the per-block groups above */
#define FD_DB_STRIDE 4
#define FD_LB_STRIDE 16

#endif

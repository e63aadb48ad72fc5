// icw_internal.h -- structures shared by the host side (icw_api.cu) and the kernels.
#pragma once
#include <cstdint>
#include "../../include/icw_b200.h"

#define ICW_PI          3.1415926535897932384626433832795029   // reference src/in_cwave.h:148
#define ICW_TWO_PI      (2.0 * ICW_PI)
#define ICW_INV_TWO_PI  (1.0 / ICW_TWO_PI)
#define ICW_SQRT2       1.4142135623730950488016887242097       // reference src/adv_modulator.c:43
#define ICW_RSQRT2      (1.0 / ICW_SQRT2)
#define ICW_SQRT6       2.4494897427831780981972840747059       // reference src/sound_render.c:50
#define ICW_HZ_SCALE    1000u                                   // reference src/in_cwave.h:162
#define ICW_SILENCE_DB  (-555.0)                                // reference src/sound_render.h:103
#define ICW_FMT_INTERNAL_F64 8     // not part of the ABI: real doubles, used by the Hilbert leaf
#define ICW_SHAPE_GENERIC      0    // interpreted DSP list
#define ICW_SHAPE_MASTER       1    // Master(In) alone: the reference's default list
#define ICW_SHAPE_SHIFT_MASTER 2    // Shift(In -> X), Master(X): no exchange, no I/Q swap
#define ICW_MT_N        624
#define ICW_MT_M        397

namespace icw {

// a DSP-list node with everything the host can fold ahead of time already folded, using the
// same IEEE operations the reference would perform per frame (so the folding is invisible)
struct DevNode {
    int32_t  mode;
    uint32_t inputs_mask;
    int32_t  xch_mode;
    int32_t  l_iq_invert, r_iq_invert;
    int32_t  n_out;
    int32_t  l_tout, r_tout;
    int32_t  l_on, r_on;
    int32_t  l_neg, r_neg;          // shift: fr_shift < 0 (adv_modulator.c:528-541)
    double   l_gain, r_gain;
    double   l_f, r_f;              // |frequency|, scaled to mHz grid when is_frmod_scaled (:36-38)
    double   l_ph0, r_ph0;          // pm: phase * PI          (:572)
    double   l_lvlpi, r_lvlpi;      // pm: level * PI          (:572)
    double   l_angle, r_angle;      // pm: angle
    // the interpreter's short cut (run_graph): a node with 1..3 inputs whose exchange mode is not (L+R)/2 takes its four
    // values straight out of the bus in the order XCH and the I/Q inversions would leave them -- the sums are per component,
    // so moving the components before or after the sum is the same arithmetic
    int32_t  n_in;                  // number of inputs when the short cut applies (1..3), else 0
    int32_t  pad_[3];
    int4     in_off[3];             // [input] .x .y .z .w: byte offset into the thread's bus[ICW_N_PLUGS][4] of the component that
                                    // ends up as value 0..3 (plugs in ascending order); one 128-bit constant load an input
};

// quantiser constants (reference sound_render_recalc, src/sound_render.c:499-551)
struct DevRender {
    double  dth_mul, hi, lo, inv_hi, norm_mul, round_off;
    int32_t neg_delta, shift, bytes;    // bytes per channel sample: 2 or 3
    int32_t render_type;
    int32_t words_per_sample;           // MT words per channel sample: 0/2/4/2/24
    // noise shaping (reference src/sound_render.c:403-489): 0 = flat, 1 = FIR, 2 = IIR of ns_order
    int32_t ns_kind, ns_order;
    double  ns_coef[2 * ICW_NS_MAX_TAPS];
};

struct DevChain {
    int32_t  fmt, n_channels, chan_bytes, frame_bytes, out_frame_bytes;
    int32_t  is_complex;
    int32_t  aligned;                   // per call: base pointer and row stride keep samples naturally aligned
    int32_t  is_frmod_scaled;
    int32_t  bypass, n_nodes;
    int32_t  shape;                     // ICW_SHAPE_*: DSP lists common enough to get straight-line code
    int32_t  filter_no, hb_ord, is_kahan, reject_flag;
    int32_t  fp_check;                  // the FP-exception-checked twins (reference src/fp_check.c:52-99)
    int32_t  feedback;                  // bit 0: a node reads a plug written later in the list: the previous frame's value,
                                        // so the DSP list is serial in time (reference src/adv_modulator.c:634-751);
                                        // bit 1 (per call): the dither draws of this call are taken one after the other with the
                                        // reference's rejection loop (replay of a frame that met it, icw_api.cu)
    int64_t  n_samples, n_fade_in, n_fade_out;
    uint64_t scale_sr;                  // sample_rate * 1000 (scaled) or 0
    double   osc_div, osc_rdiv;         // divisor of the oscillator phase and RN(1/divisor)
    double   hb_fb[ICW_MAX_ORD];        // -a[j+1]/a0        (src/hblpf.c:853)
    double   hb_ff[ICW_MAX_ORD];        //  b[j+1]/a0        (src/hblpf.c:854)
    double   hb_d0;                     //  b0/a0            (src/hblpf.c:850)
    DevRender render;
    DevNode  nodes[ICW_MAX_NODES];
};

// device-resident per-stream state: icw_stream_state plus what only the kernels need
struct DevStream {
    uint64_t n_frame;
    int64_t  pos;
    double   hb[2][2][ICW_MAX_ORD];
    unsigned long long hb_rejects[2][2];
    uint32_t quad[2];
    uint32_t mt_seed[2];
    uint64_t mt_drawn[2];
    double   prev_rnd[2];
    uint32_t clips[2];
    double   peak[2];
    double   bus[ICW_N_PLUGS][4];
    unsigned long long mt_redraws;
    double   prev_rnd_next[2];          // written by the last frame of a call, committed by advance
    uint32_t hb_basis, pad0;            // 0 = hb[] is the DF-II delay line, 1 = modal states
    double   ns_e[2][ICW_NS_MAX_TAPS], ns_o[2][ICW_NS_MAX_TAPS], ns_prev_err[2];   // age-ordered shaper memory
    // FP_EXCEPT_STATS x 4: [hilbert L, hilbert R, render L, render R][total, snan, qnan, ninf, nden, pden, pinf]
    uint32_t fp_cnt[4][7];
    // the frame-serial replay of a frame whose dither draw ran into the generator's rejection loop (mt_jrnd.c:249-253): the words
    // each channel's generator consumed
    uint32_t serial_used[2];
    uint32_t pad_fp[2];
};

}  // namespace icw

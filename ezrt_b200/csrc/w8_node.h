// w8_node.h -- the 8-wide, 8-bit-quantised node of the device's acceleration tree ("W8").
// Shared by the host builder (accel_w8.cpp), the kernels (device_functions.cuh) and the CPU model of the
// traversal (tools/w8_model.cpp), so that all three use the same layout and the same decode arithmetic.
//
// The reference's hitBVH (P5/fsh:254-306) walks a binary tree of 48-byte nodes with three dependent fetches
// per step.  The accel policy only has to find the globally closest accepted triangle (capi.cu decides by the
// deferral rule whether that is also the shader's answer), so its tree is free in shape, order and box
// precision as long as every child box is a SUPERSET of the exact (2*delta-inflated) box of its sub-tree:
//
//   record = 24 words = 96 bytes, 32-byte aligned, read with three 256-bit loads
//     w0..2   origin.xyz (float)      one quantisation step below the lowest child plane of the node
//     w3..5   scale.xyz  (float)      a power of two per axis, >= W8 minimum step of the scene (fp error bound below)
//     w6      child_base              node index of the first inner child; inner children are numbered consecutively
//                                     in slot order: child of slot s = child_base + popc(imask & ((1 << s) - 1))
//     w7      tri_base                first triangle (accel order) of the node's leaf children, consecutive in slot order
//     w8..13  qlo_x[8] qlo_y[8] qlo_z[8]   low planes of the eight slots, one byte each (slot s = byte s)
//     w14,15  meta[8]                 leaf slot: (count << 5) | offset of its first triangle from tri_base (count 1..4,
//                                     offset 0..28); inner or empty slot: 0
//     w16..21 qhi_x[8] qhi_y[8] qhi_z[8]   high planes
//     w22     imask                   bit s: slot s holds an inner child
//     w23     unused (0)
//   plane value = origin + q * scale; an empty slot has qlo = 255, qhi = 0 (inverted, never hit).
//
// Decode (one FMA per plane, same formula on host model and device):
//     B = scale * inv_d            (exact: scale is a power of two)
//     A = fma(-2^15, B, (origin - o) * inv_d)
//     f = as_float(0x47000000 | q << 8) = 2^15 + q        (one PRMT on the device: the byte goes to mantissa bits 8..15)
//     t = fma(f, B, A)             = (origin + q*scale - o) * inv_d up to rounding (the 2^15 * B terms cancel exactly)
// Conservativeness: the builder stores floor(x - W8_SLACK_STEPS) for low planes and ceil(x + W8_SLACK_STEPS) for high
// planes (x = exact plane in steps).  Rounding error of the decode: A is rounded once at magnitude <= 2^15 * B + |t|, i.e.
// 2^-9 step + 2^-24 |t|; (origin - o) * inv_d carries 2 * 2^-24 * |origin - o| * |inv_d|; the final FMA 2^-24 |t|.  The
// builder keeps scale >= W8_MIN_STEP_REL * max|coordinate| and the kernel only traces rays with
// |o| <= W8_ORIGIN_LIMIT_REL * max|coordinate| and 2^-60 <= |inv_d| <= 2^96 on this tree (all others go to the exact kernel), so
// with |plane|, |o| <= 5 max|coordinate| the total stays below 4 * 2^-24 * 5 * max|coordinate| * |inv_d| + 2^-9 step
// < 0.16 step + 0.002 step < W8_SLACK_STEPS (0.25).
//
// Slot order ("octant order", after Ylitie, Karras, Laine 2017): the builder places a child in the slot whose
// corner direction (bit a of the slot index set = towards +axis_a) matches the child's offset from the node
// centre best; a ray visits hit slots in descending (slot ^ near_mask), near_mask bit a = 1 iff d_a >= 0, so
// children on the side the ray comes from go first, without sorting distances.  The bit significance of the
// axes (which axis decides first) is a per-scene permutation: the axis of largest scene extent is bit 2.
#ifndef EZRT_W8_NODE_H
#define EZRT_W8_NODE_H

#include <stdint.h>

#define W8_NODE_WORDS 24
#define W8_NODE_BYTES 96
#define W8_MAX_LEAF_TRIS 4            // triangles per leaf slot (meta count field)
#define W8_MAX_NODE_TRIS 32           // triangles of all leaf slots of one node (bits of the triangle mask)
#define W8_SLACK_STEPS 0.25                  // outward slack of the stored planes, in quantisation steps
#define W8_DECODE_BIAS 32768.0f              // 2^15: f = as_float(W8_DECODE_BITS | q << 8) = 2^15 + q
#define W8_DECODE_BITS 0x47000000u
#define W8_MIN_STEP_REL 7.62939453125e-06f   // 2^-17: smallest quantisation step relative to max |coordinate|
#define W8_ORIGIN_LIMIT_REL 4.0f
#define W8_INV_LIMIT 7.9228162514264338e28f  // 2^96: rays with a larger |1/d_a| go to the exact kernel (no overflow in the decode)
#define W8_INV_MIN 8.6736173798840355e-19f   // 2^-60: ... or a smaller one (no underflow)

#define W8_LOCAL_STACK 48                    // stack entries beyond the shared-memory part (local memory)

#define W8_W_ORIGIN 0
#define W8_W_SCALE 3
#define W8_W_CHILD_BASE 6
#define W8_W_TRI_BASE 7
#define W8_W_QLO 8
#define W8_W_META 14
#define W8_W_QHI 16
#define W8_W_IMASK 22

#endif

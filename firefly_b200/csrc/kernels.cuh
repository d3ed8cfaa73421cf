// kernels.cuh -- launch interface of the sm_100a tape interpreter (kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

#include "tape.hpp"

namespace ffgpu {

struct EvalParams {
  const Instr* tape;          // device, resolved for prime p, terminated by two OP_HALT
  const uint64_t* values;     // device, [n_vars][in_stride], canonical residues
  uint64_t* out;              // device, [n_out][out_stride]
  uint64_t n_points;
  uint64_t in_stride;
  uint64_t out_stride;
  uint64_t p;                 // the prime
  uint32_t c;                 // p = 2^63 - c ...
  uint32_t d;                 // ... and 2c, if the prime has the form of FireFly's table (generic == 0)
  uint32_t generic;           // any other odd prime below 2^63: Barrett arithmetic on canonical residues
  uint32_t pshift;            // ... p << pshift is in [2^62, 2^63)
  uint64_t pinv;              // ... floor((2^126 - 1) / (p << pshift))
  uint32_t n_slots;           // shared-memory slots per probe: spill slots from 0 up, temp slots from n_slots-1 down
  uint32_t n_instr;           // tape length without the terminators
  uint32_t n_base;            // first variable index that is a shared row (n_vars), 0xffffffff if the collection has none
  uint32_t n_vars;            // rows of `values`; the tape's variables n_vars, n_vars + 1, ... are the shared rows:
  const uint64_t* table;      // device, [n_table_rows][table_stride], written by the collection's pre-pass tape
  uint64_t table_stride;      //   (compiler.cpp, build_table) at the same points; NULL if the collection has none
  const uint32_t* chunk_bounds;  // device, n_chunks + 1 tape offsets at function boundaries (NULL: one chunk)
};

// probes per thread / threads per block actually used for a launch
struct LaunchShape {
  int k;          // 1, 2, 4 or 8
  int threads;    // 128, 256, 512 or 1024
  unsigned blocks;     // point tiles (grid.x)
  unsigned chunks;     // tape chunks (grid.y): blocks of a tile split the collection's functions
  size_t smem;
  bool lean;           // kernel without the generic instructions, 64 registers at k = 8 (kernels.cu)
};

LaunchShape choose_shape(uint64_t n_points, uint32_t n_slots, int sm_count, int force_k, int force_threads,
                         unsigned max_chunks, bool lean_ok, bool generic_field = false);
// p = 2^63 - c with 0 < c < 2^15: the tuned arithmetic applies
bool is_table_form(uint64_t p);
// Barrett constants of any modulus below 2^63 (EvalParams::pinv, pshift)
void barrett_constants(uint64_t p, uint64_t* pinv, uint32_t* shift);
cudaError_t launch_eval(const EvalParams& prm, const LaunchShape& shape, cudaStream_t stream);

// synthetic probe points, identical on host (bench/tests) and device:
//   x[v][j] = 1 + splitmix64(seed + v * n_points + j) mod (p - 1)        in [1, p-1]
cudaError_t launch_fill_points(uint64_t* values, uint64_t n_vars, uint64_t n_points, uint64_t stride,
                               uint64_t first_point, uint64_t total_points, uint64_t seed, uint64_t p,
                               cudaStream_t stream);
// probe points on a line through the anchor: values[v][j] = a[v] * t[j] + s[v] mod p (canonical); what
// Reconstructor::queue_probes builds per point on the host (reference: Reconstructor.hpp:2946-2960)
cudaError_t launch_line_points(uint64_t* values, const uint64_t* t, const uint64_t* a, const uint64_t* s,
                               uint32_t n_vars, uint64_t n_points, uint64_t stride, uint64_t p, cudaStream_t stream);
// out[f] = (sum of column f mod p, xor of column f) over n_points -- compact fingerprints of big outputs
cudaError_t launch_checksum(const uint64_t* out, uint64_t n_out, uint64_t n_points, uint64_t stride, uint64_t p,
                            uint64_t* sums /* [n_out][2] */, cudaStream_t stream);
// field-arithmetic self test: r[i] = op(a[i], b[i]) for the device primitives (tests only call this)
cudaError_t launch_field_ops(int op, const uint64_t* a, const uint64_t* b, uint64_t* r, uint64_t n, uint64_t p,
                             cudaStream_t stream);

}  // namespace ffgpu

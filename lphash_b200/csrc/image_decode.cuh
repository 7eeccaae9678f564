// Device-side loader: builds the flat device image (device_image.h) from the pieces of a serialized index
// (ImagePlan, lph_image.h) on the GPU - the reference's read-side primitives evaluated for every bucket at once:
//   pthash::compact_vector::access      pthash/include/encoders/compact_vector.hpp:229-234   (pilot ranks / dictionaries)
//   Elias-Fano access via select        include/ef_sequence.hpp:77-81, pthash ef_sequence / darray.hpp:50-76
//   rs_bit_vector::rank                 include/rs_bit_vector.hpp:32-43
//   quartet_wtree::rank_of              src/quartet_wtree.cpp:84-99
//   the per-bucket part of mphf::query  src/partitioned_mphf.cpp:292-339 (mphf_alt: src/unpartitioned_mphf.cpp:191-206)
// Same results, byte for byte, as the host decode of lph_image.cpp (tests compare the two arenas).
#pragma once
#include <cuda_runtime.h>

#include "lph_image.h"

namespace lphb {

// d_arena: plan.arena_bytes bytes on the current device.  Synchronous; throws FormatError for the inconsistencies
// only decoding reveals (pilot rank outside its dictionary, free slot outside the table, ...), CudaError otherwise.
// Sets *collision_base for the partitioned form (needs one decoded value).
void decode_image_on_device(ImagePlan const& plan, void* d_arena, uint64_t* collision_base);

// The same for a plan of ImageBuilder::plan_phf: pilot_hash and free32 of plan.img.minimizer_order.
void decode_phf_on_device(ImagePlan const& plan, void* d_arena);

DevImage rebase_image(DevImage img, const void* device_base);

}  // namespace lphb

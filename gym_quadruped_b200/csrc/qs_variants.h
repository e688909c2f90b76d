// qs_variants.h -- registry of the compiled kernel variants (one translation unit each, qs_inst_*.cu, built in parallel).
#pragma once
#include <vector>

#include "qs_kernel.cuh"

namespace qs {

// generic variants (FEAT = 0): step + reset + forward (+ the stand-alone ray-cast kernel on the condim-3 ones)
VariantInfo variant_f3();
VariantInfo variant_f6();
VariantInfo variant_d3();
VariantInfo variant_d6();
// specialised fp32 step kernels for the BASELINE configurations
VariantInfo variant_f3_pyr_flat_mesh();   // configs[1]: mini_cheetah / flat                     (pyramidal, sphere + mesh geoms, no limits)
VariantInfo variant_f3_pyr_hfield_prim(); // configs[2]: aliengo / perlin + height map           (pyramidal, sphere + capsule + box geoms)
VariantInfo variant_f6_ell_boxes_prim();  // configs[3]: go2 / random_boxes                      (elliptic, condim 6, primitives)
VariantInfo variant_f3_ell_flat_mesh();   // configs[4]: hyqreal1 / flat + IMU                   (elliptic, sphere + mesh geoms)
// flat-floor variants for the remaining robots
VariantInfo variant_f3_pyr_flat_prim();   // aliengo, hyqreal2, b2 / flat                        (pyramidal, primitives)
VariantInfo variant_f6_ell_flat_prim();   // go2, go1 / flat                                     (elliptic, condim 6, primitives)
VariantInfo variant_f6_ell_flat_mesh();   // spot / flat                                         (elliptic, condim 6, sphere + mesh geoms)

inline const std::vector<VariantInfo>& all_variants() {
  static const std::vector<VariantInfo> v = {variant_f3(), variant_f6(), variant_d3(), variant_d6(), variant_f3_pyr_flat_mesh(),
                                             variant_f3_pyr_hfield_prim(), variant_f6_ell_boxes_prim(), variant_f3_ell_flat_mesh(),
                                             variant_f3_pyr_flat_prim(), variant_f6_ell_flat_prim(), variant_f6_ell_flat_mesh()};
  return v;
}

}  // namespace qs

// qs_inst_f3_pyr_flat_prim.cu -- instantiates one kernel variant (see qs_variants.h); compiled in parallel with its siblings.
#include "qs_variants.h"

namespace qs {
VariantInfo variant_f3_pyr_flat_prim() {
  VariantInfo v = make_variant<float, 3, FEAT_PYR_FLAT_PRIM, false>("f3_pyr_flat_prim");

  return v;
}
}  // namespace qs

// qs_inst_d3.cu -- instantiates one kernel variant (see qs_variants.h); compiled in parallel with its siblings.
#include "qs_variants.h"

namespace qs {
VariantInfo variant_d3() {
  VariantInfo v = make_variant<double, 3, 0, true>("d3");
  v.raycast = raycast_kernel<double>;
  return v;
}
}  // namespace qs

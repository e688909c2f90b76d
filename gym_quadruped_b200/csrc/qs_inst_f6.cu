// qs_inst_f6.cu -- instantiates one kernel variant (see qs_variants.h); compiled in parallel with its siblings.
#include "qs_variants.h"

namespace qs {
VariantInfo variant_f6() {
  VariantInfo v = make_variant<float, 6, 0, true>("f6");

  return v;
}
}  // namespace qs

// dist.cu -- multi-GPU sharding (one process per GPU).  Placeholder until the NCCL
// qubit-swap path lands: every entry point reports QCB_EUNSUPPORTED.
#include "qcb_internal.hpp"

namespace qcb {
int dist_state_canonicalize(qcb_state *) { return set_error(QCB_EUNSUPPORTED, "sharded states are not available yet"); }
int dist_make_progress(qcb_state *, const std::vector<GateOp> &, const std::vector<char> &)
{
    return set_error(QCB_EUNSUPPORTED, "sharded states are not available yet");
}
} // namespace qcb

extern "C" {
int qcb_dist_unique_id(void *) { return qcb::set_error(QCB_EUNSUPPORTED, "sharded states are not available yet"); }
int qcb_dist_init(int, int, const void *) { return qcb::set_error(QCB_EUNSUPPORTED, "sharded states are not available yet"); }
int qcb_dist_finalize(void) { return QCB_OK; }
}

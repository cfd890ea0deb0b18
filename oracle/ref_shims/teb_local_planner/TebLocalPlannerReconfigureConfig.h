#ifndef REF_SHIM_RECONF
#define REF_SHIM_RECONF
namespace teb_local_planner { class TebLocalPlannerReconfigureConfig {}; }
#endif

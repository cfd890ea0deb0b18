#include <teb_local_planner/TrajectoryMsg.h>

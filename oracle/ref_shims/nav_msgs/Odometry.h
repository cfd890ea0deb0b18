#include <nav_msgs/Path.h>

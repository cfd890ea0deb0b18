#include <g2o/core/sparse_optimizer.h>

#include <ros/console.h>

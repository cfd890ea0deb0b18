#ifndef REF_SHIM_GM_POINT
#define REF_SHIM_GM_POINT
#include <vector>
#include <string>
namespace std_msgs { struct Header { std::string frame_id; }; struct ColorRGBA { float r, g, b, a; ColorRGBA() : r(0), g(0), b(0), a(0) {} }; }
namespace geometry_msgs {
struct Point { double x, y, z; Point() : x(0), y(0), z(0) {} };
struct Point32 { float x, y, z; Point32() : x(0), y(0), z(0) {} };
struct Vector3 { double x, y, z; Vector3() : x(0), y(0), z(0) {} };
struct Quaternion { double x, y, z, w; Quaternion() : x(0), y(0), z(0), w(1) {} };
struct Pose { Point position; Quaternion orientation; };
struct PoseStamped { std_msgs::Header header; Pose pose; };
struct Twist { Vector3 linear, angular; };
struct TwistWithCovariance { Twist twist; double covariance[36]; };
struct QuaternionStamped { std_msgs::Header header; Quaternion quaternion; };
struct Polygon { std::vector<Point32> points; };
struct TwistStamped { std_msgs::Header header; Twist twist; };
struct PoseArray { std_msgs::Header header; std::vector<Pose> poses; };
}
#endif

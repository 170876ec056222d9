// b2_compat.hpp -- the types the sensor wrappers need from ROS 2 and MuJoCo.
//
// With -DB2_HAVE_ROS2 (and MuJoCo's headers on the include path) this header just includes the
// real ones, and mujoco_lidar.cpp / mujoco_cameras.cpp compile into the reference's plugin as
// drop-in replacements of ref:src/mujoco_lidar.cpp and ref:src/mujoco_cameras.cpp.
//
// Neither ROS 2 nor MuJoCo exists in this build image, so by default it
// declares minimal stand-ins with the SAME names, field names and layouts the wrappers touch:
//   sensor_msgs::msg::{LaserScan, Image, CameraInfo, PointCloud2, PointField}, std_msgs Header,
//   hardware_interface::{HardwareInfo, ComponentInfo}, rclcpp::{Node, Publisher, Rate, ok},
//   mjModel / mjData restricted to the fields of SURVEY.md A.1.
// The stand-in publisher keeps the last message so tests can read what would have gone on the wire.
#ifndef B2_COMPAT_HPP
#define B2_COMPAT_HPP

#ifdef B2_HAVE_ROS2
#include <mujoco/mujoco.h>

#include <hardware_interface/hardware_info.hpp>
#include <rclcpp/rclcpp.hpp>
#include <sensor_msgs/image_encodings.hpp>
#include <sensor_msgs/msg/camera_info.hpp>
#include <sensor_msgs/msg/image.hpp>
#include <sensor_msgs/msg/laser_scan.hpp>
#include <sensor_msgs/msg/point_cloud2.hpp>
#else

#include <array>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

// ---- MuJoCo subset -----------------------------------------------------------------------------
typedef double mjtNum;
enum mjtSensor_ { mjSENS_RANGEFINDER = 7 };
namespace mjtSensor { constexpr int mjSENS_RANGEFINDER = ::mjSENS_RANGEFINDER; }
enum mjtObj_ { mjOBJ_SENSOR = 20 };

struct mjVisualMap_ {
  float znear = 0.01f, zfar = 50.0f;
};
struct mjVisual_ {
  mjVisualMap_ map;
};
struct mjStatistic_ {
  mjtNum extent = 1.0;
};

struct mjModel {
  int ngeom = 0, nbody = 0, nmesh = 0, nmat = 0, nsite = 0, ncam = 0, nsensor = 0, nsensordata = 0;
  // geoms / bodies / meshes (ray_eliminate, mju_rayGeom, mj_rayMesh)
  int* geom_type = nullptr;
  int* geom_bodyid = nullptr;
  int* geom_group = nullptr;
  int* geom_matid = nullptr;
  int* geom_dataid = nullptr;
  mjtNum* geom_size = nullptr;
  float* geom_rgba = nullptr;
  float* mat_rgba = nullptr;
  int* body_weldid = nullptr;
  int* mesh_vertadr = nullptr;
  int* mesh_vertnum = nullptr;
  int* mesh_faceadr = nullptr;
  int* mesh_facenum = nullptr;
  float* mesh_vert = nullptr;
  int* mesh_face = nullptr;
  // sites / sensors / cameras
  int* site_bodyid = nullptr;
  mjtNum* site_pos = nullptr;   // [nsite*3] in the body frame
  mjtNum* site_quat = nullptr;  // [nsite*4] in the body frame
  int* sensor_type = nullptr;
  int* sensor_objid = nullptr;
  int* sensor_adr = nullptr;
  mjtNum* sensor_cutoff = nullptr;
  int* name_sensoradr = nullptr;
  int* name_camadr = nullptr;
  char* names = nullptr;
  int* cam_resolution = nullptr;  // [ncam*2]
  mjtNum* cam_fovy = nullptr;
  mjVisual_ vis;
  mjStatistic_ stat;
};

struct mjData {
  mjtNum time = 0;
  mjtNum* geom_xpos = nullptr;  // [ngeom*3]
  mjtNum* geom_xmat = nullptr;  // [ngeom*9]
  mjtNum* xpos = nullptr;       // [nbody*3]
  mjtNum* xmat = nullptr;       // [nbody*9]
  mjtNum* cam_xpos = nullptr;   // [ncam*3]
  mjtNum* cam_xmat = nullptr;   // [ncam*9]
  mjtNum* sensordata = nullptr;
};

inline const char* mj_id2name(const mjModel* m, int type, int id) {
  if (type == mjOBJ_SENSOR && m->names && m->name_sensoradr && id >= 0 && id < m->nsensor)
    return m->names + m->name_sensoradr[id];
  return nullptr;
}

// ---- ROS 2 message subset -----------------------------------------------------------------------
namespace builtin_interfaces::msg {
struct Time {
  int32_t sec = 0;
  uint32_t nanosec = 0;
};
}  // namespace builtin_interfaces::msg

namespace std_msgs::msg {
struct Header {
  builtin_interfaces::msg::Time stamp;
  std::string frame_id;
};
}  // namespace std_msgs::msg

namespace sensor_msgs {
namespace image_encodings {
constexpr const char* RGB8 = "rgb8";
constexpr const char* TYPE_32FC1 = "32FC1";
}  // namespace image_encodings
namespace msg {
struct LaserScan {
  std_msgs::msg::Header header;
  float angle_min = 0, angle_max = 0, angle_increment = 0, time_increment = 0, scan_time = 0;
  float range_min = 0, range_max = 0;
  std::vector<float> ranges, intensities;
};
struct Image {
  std_msgs::msg::Header header;
  uint32_t height = 0, width = 0;
  std::string encoding;
  uint8_t is_bigendian = 0;
  uint32_t step = 0;
  std::vector<uint8_t> data;
};
struct CameraInfo {
  std_msgs::msg::Header header;
  uint32_t height = 0, width = 0;
  std::string distortion_model;
  std::vector<double> d;
  std::array<double, 9> k{}, r{};
  std::array<double, 12> p{};
};
struct PointField {
  static constexpr uint8_t FLOAT32 = 7;
  std::string name;
  uint32_t offset = 0;
  uint8_t datatype = 0;
  uint32_t count = 0;
};
struct PointCloud2 {
  std_msgs::msg::Header header;
  uint32_t height = 0, width = 0;
  std::vector<PointField> fields;
  bool is_bigendian = false;
  uint32_t point_step = 0, row_step = 0;
  std::vector<uint8_t> data;
  bool is_dense = false;
};
}  // namespace msg
}  // namespace sensor_msgs

// ---- ros2_control subset --------------------------------------------------------------------------
namespace hardware_interface {
struct ComponentInfo {
  std::string name;
  std::unordered_map<std::string, std::string> parameters;
};
struct HardwareInfo {
  std::vector<ComponentInfo> sensors;
};
}  // namespace hardware_interface

// ---- rclcpp subset -----------------------------------------------------------------------------------
namespace rclcpp {
inline bool ok() { return true; }

template <class MsgT>
class Publisher {
public:
  using SharedPtr = std::shared_ptr<Publisher<MsgT>>;
  explicit Publisher(std::string topic) : topic_(std::move(topic)) {}
  void publish(const MsgT& msg) {
    last_ = msg;
    ++count_;
  }
  const std::string& get_topic_name() const { return topic_; }
  const MsgT& last() const { return last_; }
  size_t count() const { return count_; }

private:
  std::string topic_;
  MsgT last_;
  size_t count_ = 0;
};

// logging stand-in: messages are kept (tests read them) and echoed to stderr when B2_COMPAT_LOG is set
struct Logger {
  std::shared_ptr<std::vector<std::string>> lines = std::make_shared<std::vector<std::string>>();
  void log(const char* level, const std::string& text) const {
    lines->push_back(std::string(level) + ": " + text);
    if (std::getenv("B2_COMPAT_LOG")) std::fprintf(stderr, "[%s] %s\n", level, text.c_str());
  }
};

class Node {
public:
  using SharedPtr = std::shared_ptr<Node>;
  const Logger& get_logger() const { return logger_; }

private:
  Logger logger_;

public:
  template <class MsgT>
  typename Publisher<MsgT>::SharedPtr create_publisher(const std::string& topic, size_t /*depth*/) {
    return std::make_shared<Publisher<MsgT>>(topic);
  }
  builtin_interfaces::msg::Time now() const {
    const auto t = std::chrono::steady_clock::now().time_since_epoch();
    const auto ns = std::chrono::duration_cast<std::chrono::nanoseconds>(t).count();
    builtin_interfaces::msg::Time out;
    out.sec = (int32_t)(ns / 1000000000LL);
    out.nanosec = (uint32_t)(ns % 1000000000LL);
    return out;
  }
};

class Rate {
public:
  explicit Rate(double hz) : period_(hz > 0 ? 1.0 / hz : 0.0) {}
  void sleep() { std::this_thread::sleep_for(std::chrono::duration<double>(period_)); }

private:
  double period_;
};
}  // namespace rclcpp

// printf-style, like the real macros; the stream variants take `a << b`
#define B2_COMPAT_LOGF(level, logger, ...)              \
  do {                                                  \
    char b2_buf_[512];                                  \
    std::snprintf(b2_buf_, sizeof(b2_buf_), __VA_ARGS__); \
    (logger).log(level, b2_buf_);                       \
  } while (0)
#define RCLCPP_INFO(logger, ...) B2_COMPAT_LOGF("INFO", logger, __VA_ARGS__)
#define RCLCPP_WARN(logger, ...) B2_COMPAT_LOGF("WARN", logger, __VA_ARGS__)
#define RCLCPP_ERROR(logger, ...) B2_COMPAT_LOGF("ERROR", logger, __VA_ARGS__)

#endif  // B2_HAVE_ROS2
#endif

// mujoco_lidar.hpp -- B200 drop-in for ref:include/mujoco_ros2_control/mujoco_lidar.hpp.
//
// Interface-compatible re-implementation: class / method / field names mirror the reference header
// (yassine-cherni/mujoco_ros2_simulation, Copyright (c) 2025 United States Government as represented
// by the Administrator of NASA, Apache License 2.0) because the plugin that constructs this class
// must not notice the swap; the GPU members and update() are new.
//
// Same class, same public methods (register_lidar / init / close), same URDF sensor parameters
// (frame_name, min_angle, max_angle, angle_increment, laserscan_topic, range_min, range_max) and
// the same LaserScan message (ref:src/mujoco_lidar.cpp:54-112).  What changes is where the ranges
// come from: the reference copies them out of mjData::sensordata, where MuJoCo's rangefinder pass
// left them (one brute-force mj_ray per ray inside every mj_step); here update() casts the whole
// fan on the GPU through the C ABI (b2_lidar_cast), with the float cast and the range filter of
// ref:src/mujoco_lidar.cpp:262,269-271 fused into the kernel's epilogue.
#ifndef B2_MUJOCO_LIDAR_HPP
#define B2_MUJOCO_LIDAR_HPP

#include <atomic>
#include <mutex>
#include <optional>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include "../../include/b2sense.h"
#include "b2_compat.hpp"
#include "scene_mirror.hpp"

namespace mujoco_ros2_control
{

struct LidarData
{
  std::string name;
  std::string frame_name;
  int num_rangefinders;
  double min_angle;
  double max_angle;
  double angle_increment;
  double range_min;
  double range_max;

  // Maps the index of the rangefinder to the index of the mujoco rangefinder's data
  // (kept for parity with the reference; the GPU path does not read sensordata).
  std::vector<int> sensor_indexes;

  // GPU path: ray i of the fan in the frame of the body that carries the rangefinder sites
  std::vector<double> local_dir;    // [num_rangefinders*3] site +Z axes (engine_sensor.c)
  std::vector<uint8_t> have_ray;    // a rangefinder sensor was found for index i
  int bodyid{ -1 };                 // site_bodyid: the rangefinder's bodyexclude
  double site_pos[3]{ 0, 0, 0 };    // common site position in that body
  double cutoff{ 0.0 };             // sensor_cutoff (one per lidar: register_lidar checks)
  std::vector<double> ray_cutoff;   // sensor_cutoff of each replica
  b2_lidar* engine{ nullptr };

  // For message publishing
  std::string laserscan_topic;
  sensor_msgs::msg::LaserScan laser_scan_msg;
  rclcpp::Publisher<sensor_msgs::msg::LaserScan>::SharedPtr scan_pub;
};

std::pair<std::string, int> parse_lidar_name(const std::string& sensor_name);
std::optional<LidarData> get_lidar_data(const hardware_interface::HardwareInfo& hardware_info,
                                        const std::string& name);

/**
 * @brief Wraps MuJoCo rangefinder fans for publishing ROS 2 LaserScan messages, cast on the GPU.
 */
class MujocoLidar
{
public:
  /**
   * Same arguments as the reference constructor plus the scene mirror the interface owns.
   */
  explicit MujocoLidar(rclcpp::Node::SharedPtr& node, std::recursive_mutex* sim_mutex, mjData* mujoco_data,
                       mjModel* mujoco_model, double lidar_publish_rate, SceneMirror* mirror = nullptr);
  ~MujocoLidar();

  /** Starts the lidar processing thread in the background (creates the GPU objects first). */
  void init();
  /** Stops the lidar processing thread and releases the GPU objects. */
  void close();
  /** Parses lidar information from the mujoco model. Returns true if successful. */
  bool register_lidar(const hardware_interface::HardwareInfo& hardware_info);

  // One publish cycle (what the thread runs at lidar_publish_rate); public for tests.
  bool create_engines();
  void update();
  const std::vector<LidarData>& lidars() const
  {
    return lidar_sensors_;
  }
  std::string last_error() const
  {
    std::lock_guard<std::mutex> g(error_mutex_);
    return last_error_;
  }
  // When the physics thread does not call SceneMirror::refresh after mj_step (shim v0), update()
  // refreshes the mirror itself while it holds the sim mutex.
  void set_refresh_in_update(bool v)
  {
    refresh_in_update_ = v;
  }

private:
  void update_loop();
  void report_failure(const std::string& what);

  rclcpp::Node::SharedPtr node_;
  std::recursive_mutex* sim_mutex_{ nullptr };
  mjData* mj_data_;
  mjModel* mj_model_;
  double lidar_publish_rate_;
  SceneMirror* mirror_;
  bool refresh_in_update_{ true };

  std::vector<LidarData> lidar_sensors_;
  std::vector<double> pose_scratch_;  // [12 per lidar] sensor frames of the cycle in flight
  std::thread rendering_thread_;
  std::atomic_bool publish_lidar_{ false };
  mutable std::mutex error_mutex_;
  std::string last_error_;
  unsigned failures_{ 0 };
};

}  // namespace mujoco_ros2_control
#endif

// mujoco_cameras.hpp -- B200 drop-in for ref:include/mujoco_ros2_control/mujoco_cameras.hpp.
//
// Interface-compatible re-implementation: class / method / field names mirror the reference header
// (yassine-cherni/mujoco_ros2_simulation, Copyright (c) 2025 United States Government as represented
// by the Administrator of NASA, Apache License 2.0) because the plugin that constructs this class
// must not notice the swap; the PointCloud2 members, the GPU members and update() are new.
//
// Same class and public methods (register_cameras / init / close), same URDF sensor parameters
// (frame_name, info_topic, image_topic, depth_topic -- all-or-nothing, ref:src/mujoco_cameras.cpp:63-66),
// same Image / CameraInfo fields (:91-124, including the all-zero CameraInfo.r quirk).  Step 3 of
// update() -- float32 depth linearisation + vertical flip (:222-240) -- runs on the GPU fused with
// the PointCloud2 projection (b2_depth_to_cloud).  New: a PointCloud2 publisher per camera on
// <depth_topic>/points (depth_image_proc::point_cloud_xyz semantics).
//
// The raw GL depth buffer comes from mjr_readPixels in the real plugin (render_gl_depth hook);
// without OpenGL the wrapper ray-casts it from the scene mirror (b2_camera_render_depth_device,
// SURVEY.md section 8(f) N1).
#ifndef B2_MUJOCO_CAMERAS_HPP
#define B2_MUJOCO_CAMERAS_HPP

#include <atomic>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/b2sense.h"
#include "b2_compat.hpp"
#include "scene_mirror.hpp"

namespace mujoco_ros2_control
{

struct CameraData
{
  int cam_id{ 0 };
  std::string name;
  std::string frame_name;
  std::string info_topic;
  std::string image_topic;
  std::string depth_topic;
  std::string points_topic;

  uint32_t width;
  uint32_t height;

  std::vector<uint8_t> image_buffer;  // RGB as read from GL (bottom-up)
  std::vector<float> depth_buffer;    // raw GL depth as read from GL (bottom-up, 0 near .. 1 far)

  sensor_msgs::msg::Image image;
  sensor_msgs::msg::Image depth_image;
  sensor_msgs::msg::CameraInfo camera_info;
  sensor_msgs::msg::PointCloud2 cloud;

  rclcpp::Publisher<sensor_msgs::msg::Image>::SharedPtr image_pub;
  rclcpp::Publisher<sensor_msgs::msg::Image>::SharedPtr depth_image_pub;
  rclcpp::Publisher<sensor_msgs::msg::CameraInfo>::SharedPtr camera_info_pub;
  rclcpp::Publisher<sensor_msgs::msg::PointCloud2>::SharedPtr cloud_pub;

  b2_camera* engine{ nullptr };
};

class MujocoCameras
{
public:
  explicit MujocoCameras(rclcpp::Node::SharedPtr& node, std::recursive_mutex* sim_mutex, mjData* mujoco_data,
                         mjModel* mujoco_model, double camera_publish_rate, SceneMirror* mirror = nullptr);
  ~MujocoCameras();

  void init();
  void close();
  void register_cameras(const hardware_interface::HardwareInfo& hardware_info);

  // Hook that fills camera.depth_buffer (+ image_buffer) for one camera: mjv_updateScene /
  // mjr_render / mjr_readPixels in the real plugin (ref:src/mujoco_cameras.cpp:214-218).  When
  // unset, the depth buffer is ray-cast from the scene mirror on the GPU.
  using GlReadback = std::function<bool(CameraData&)>;
  void set_gl_readback(GlReadback fn)
  {
    gl_readback_ = std::move(fn);
  }

  bool create_engines();
  void update();
  std::vector<CameraData>& cameras()
  {
    return cameras_;
  }
  std::string last_error() const
  {
    std::lock_guard<std::mutex> g(error_mutex_);
    return last_error_;
  }
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
  double camera_publish_rate_;
  SceneMirror* mirror_;
  bool refresh_in_update_{ true };
  GlReadback gl_readback_;

  std::vector<CameraData> cameras_;
  std::thread rendering_thread_;
  std::atomic_bool publish_images_{ false };
  std::vector<double> pose_scratch_;  // [12 per camera] camera frames of the cycle in flight
  mutable std::mutex error_mutex_;
  std::string last_error_;
  unsigned failures_{ 0 };
};

}  // namespace mujoco_ros2_control
#endif

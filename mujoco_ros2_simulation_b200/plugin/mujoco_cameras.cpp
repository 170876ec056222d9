// mujoco_cameras.cpp -- B200 drop-in for ref:src/mujoco_cameras.cpp (see mujoco_cameras.hpp).
#include "mujoco_cameras.hpp"

#include <cmath>
#include <cstring>
#include <optional>

namespace mujoco_ros2_control
{

namespace
{
std::optional<hardware_interface::ComponentInfo>
get_sensor_from_info(const hardware_interface::HardwareInfo& hardware_info, const std::string& name)
{
  for (const auto& sensor : hardware_info.sensors)
  {
    if (sensor.name == name)
    {
      return sensor;
    }
  }
  return std::nullopt;
}
}  // namespace

MujocoCameras::MujocoCameras(rclcpp::Node::SharedPtr& node, std::recursive_mutex* sim_mutex, mjData* mujoco_data,
                             mjModel* mujoco_model, double camera_publish_rate, SceneMirror* mirror)
  : node_(node)
  , sim_mutex_(sim_mutex)
  , mj_data_(mujoco_data)
  , mj_model_(mujoco_model)
  , camera_publish_rate_(camera_publish_rate)
  , mirror_(mirror)
{
}

MujocoCameras::~MujocoCameras()
{
  close();
}

// ref:src/mujoco_cameras.cpp:40-129 -- same defaults, same message fields
void MujocoCameras::register_cameras(const hardware_interface::HardwareInfo& hardware_info)
{
  cameras_.resize(0);
  for (auto i = 0; i < mj_model_->ncam; ++i)
  {
    const char* cam_name = mj_model_->names + mj_model_->name_camadr[i];
    const int* cam_resolution = mj_model_->cam_resolution + 2 * i;
    const mjtNum cam_fovy = mj_model_->cam_fovy[i];

    CameraData camera;
    camera.cam_id = i;
    camera.name = cam_name;
    camera.width = static_cast<uint32_t>(cam_resolution[0]);
    camera.height = static_cast<uint32_t>(cam_resolution[1]);

    const auto camera_info_maybe = get_sensor_from_info(hardware_info, cam_name);
    if (camera_info_maybe.has_value())
    {
      // all-or-nothing: .at() throws if one is missing, as in the reference (:63-66)
      const auto camera_info = camera_info_maybe.value();
      camera.frame_name = camera_info.parameters.at("frame_name");
      camera.info_topic = camera_info.parameters.at("info_topic");
      camera.image_topic = camera_info.parameters.at("image_topic");
      camera.depth_topic = camera_info.parameters.at("depth_topic");
    }
    else
    {
      camera.frame_name = camera.name + "_frame";
      camera.info_topic = camera.name + "/camera_info";
      camera.image_topic = camera.name + "/color";
      camera.depth_topic = camera.name + "/depth";
    }
    camera.points_topic = camera.depth_topic + "/points";

    camera.camera_info_pub = node_->create_publisher<sensor_msgs::msg::CameraInfo>(camera.info_topic, 1);
    camera.image_pub = node_->create_publisher<sensor_msgs::msg::Image>(camera.image_topic, 1);
    camera.depth_image_pub = node_->create_publisher<sensor_msgs::msg::Image>(camera.depth_topic, 1);
    camera.cloud_pub = node_->create_publisher<sensor_msgs::msg::PointCloud2>(camera.points_topic, 1);

    camera.image.header.frame_id = camera.frame_name;
    const auto image_size = camera.width * camera.height * 3;
    camera.image_buffer.resize(image_size);
    camera.image.data.resize(image_size);
    camera.image.width = camera.width;
    camera.image.height = camera.height;
    camera.image.step = camera.width * 3;
    camera.image.encoding = sensor_msgs::image_encodings::RGB8;

    camera.depth_image.header.frame_id = camera.frame_name;
    camera.depth_buffer.resize(camera.width * camera.height);
    camera.depth_image.data.resize(camera.width * camera.height * sizeof(float));
    camera.depth_image.width = camera.width;
    camera.depth_image.height = camera.height;
    camera.depth_image.step = camera.width * sizeof(float);
    camera.depth_image.encoding = sensor_msgs::image_encodings::TYPE_32FC1;

    camera.camera_info.header.frame_id = camera.frame_name;
    camera.camera_info.width = camera.width;
    camera.camera_info.height = camera.height;
    camera.camera_info.distortion_model = "plumb_bob";
    camera.camera_info.k.fill(0.0);
    camera.camera_info.r.fill(0.0);  // left all-zero by the reference (:115); kept
    camera.camera_info.p.fill(0.0);
    camera.camera_info.d.resize(5, 0.0);
    double focal_scaling = (1.0 / std::tan((cam_fovy * M_PI / 180.0) / 2.0)) * camera.height / 2.0;
    camera.camera_info.k[0] = camera.camera_info.p[0] = focal_scaling;
    camera.camera_info.k[2] = camera.camera_info.p[2] = static_cast<double>(camera.width) / 2.0;
    camera.camera_info.k[4] = camera.camera_info.p[5] = focal_scaling;
    camera.camera_info.k[5] = camera.camera_info.p[6] = static_cast<double>(camera.height) / 2.0;
    camera.camera_info.k[8] = camera.camera_info.p[10] = 1.0;

    // PointCloud2: organised, x/y/z FLOAT32 at 0/4/8, point_step 12, NaN for invalid depth
    camera.cloud.header.frame_id = camera.frame_name;
    camera.cloud.height = camera.height;
    camera.cloud.width = camera.width;
    camera.cloud.fields.resize(3);
    const char* names[3] = { "x", "y", "z" };
    for (int f = 0; f < 3; ++f)
    {
      camera.cloud.fields[f].name = names[f];
      camera.cloud.fields[f].offset = 4 * f;
      camera.cloud.fields[f].datatype = sensor_msgs::msg::PointField::FLOAT32;
      camera.cloud.fields[f].count = 1;
    }
    camera.cloud.is_bigendian = false;
    camera.cloud.point_step = 12;
    camera.cloud.row_step = camera.width * 12;
    camera.cloud.is_dense = false;
    camera.cloud.data.resize((size_t)camera.width * camera.height * 12);

    cameras_.push_back(camera);
  }
}

bool MujocoCameras::create_engines()
{
  for (auto& camera : cameras_)
  {
    if (camera.engine != nullptr)
    {
      continue;
    }
    b2_camera_desc d{};
    d.width = (int32_t)camera.width;
    d.height = (int32_t)camera.height;
    d.fovy = mj_model_->cam_fovy[camera.cam_id];
    d.znear = mj_model_->vis.map.znear;  // ref :222-223: near = znear * extent, far = zfar * extent
    d.zfar = mj_model_->vis.map.zfar;
    d.extent = mj_model_->stat.extent;
    d.point_step = 12;
    if (b2_camera_create(mirror_ ? mirror_->handle() : nullptr, 0, &d, &camera.engine) != B2_OK)
    {
      last_error_ = std::string("b2_camera_create: ") + b2_last_error();
      return false;
    }
  }
  return true;
}

void MujocoCameras::init()
{
  // the reference bails out when GLFW cannot start (:134-139); here the GPU objects must exist
  if (!create_engines())
  {
    publish_images_ = false;
    return;
  }
  publish_images_ = true;
  rendering_thread_ = std::thread(&MujocoCameras::update_loop, this);
}

void MujocoCameras::close()
{
  publish_images_ = false;
  if (rendering_thread_.joinable())
  {
    rendering_thread_.join();
  }
  for (auto& camera : cameras_)
  {
    if (camera.engine != nullptr)
    {
      b2_camera_destroy(camera.engine);
      camera.engine = nullptr;
    }
  }
}

void MujocoCameras::update_loop()
{
  rclcpp::Rate rate(camera_publish_rate_);
  while (rclcpp::ok() && publish_images_)
  {
    update();
    rate.sleep();
  }
}

void MujocoCameras::update()
{
  // Step 1: under the sim mutex take the camera poses (and refresh the mirror in shim v0)
  std::vector<double> pos(3 * cameras_.size()), mat(9 * cameras_.size());
  {
    std::unique_lock<std::recursive_mutex> lock(*sim_mutex_);
    if (refresh_in_update_ && mirror_ != nullptr)
    {
      mirror_->refresh(mj_data_);
    }
    for (size_t k = 0; k < cameras_.size(); ++k)
    {
      std::memcpy(&pos[3 * k], mj_data_->cam_xpos + 3 * cameras_[k].cam_id, 3 * sizeof(double));
      std::memcpy(&mat[9 * k], mj_data_->cam_xmat + 9 * cameras_[k].cam_id, 9 * sizeof(double));
    }
  }

  for (size_t k = 0; k < cameras_.size(); ++k)
  {
    auto& camera = cameras_[k];
    float* depth_out = reinterpret_cast<float*>(camera.depth_image.data.data());
    int rc;
    if (gl_readback_)
    {
      // Step 2 (reference path): GL render + readback fills depth_buffer / image_buffer;
      // Step 3 on the GPU: linearise + flip + cloud in one kernel
      if (!gl_readback_(camera))
      {
        continue;
      }
      rc = b2_depth_to_cloud(camera.engine, 1, camera.depth_buffer.data(), depth_out, camera.cloud.data.data());
      // RGB: GL origin is bottom-left, flip row by row (:243-249)
      const auto row_size = camera.width * 3;
      for (uint32_t h = 0; h < camera.height; h++)
      {
        std::memcpy(&camera.image.data[(camera.height - 1 - h) * row_size], &camera.image_buffer[h * row_size],
                    row_size);
      }
    }
    else
    {
      // no OpenGL: depth by ray casting the scene mirror, then the same fused step 3
      rc = b2_camera_render_to_cloud(camera.engine, 1, nullptr, &pos[3 * k], &mat[9 * k], camera.depth_buffer.data(),
                                     depth_out, camera.cloud.data.data());
    }
    if (rc != B2_OK)
    {
      last_error_ = std::string("camera update: ") + b2_last_error();
      return;
    }
  }

  // Step 4: publish
  for (auto& camera : cameras_)
  {
    const auto time = node_->now();
    camera.image.header.stamp = time;
    camera.depth_image.header.stamp = time;
    camera.camera_info.header.stamp = time;
    camera.cloud.header.stamp = time;
    camera.image_pub->publish(camera.image);
    camera.depth_image_pub->publish(camera.depth_image);
    camera.camera_info_pub->publish(camera.camera_info);
    camera.cloud_pub->publish(camera.cloud);
  }
}

}  // namespace mujoco_ros2_control

// mujoco_cameras.cpp -- B200 drop-in for ref:src/mujoco_cameras.cpp (see mujoco_cameras.hpp).
//
// Interface-compatible re-implementation.  Class and method names, the URDF sensor parameter names,
// topic defaults and the Image / CameraInfo field values follow the reference file
// (yassine-cherni/mujoco_ros2_simulation, src/mujoco_cameras.cpp, Copyright (c) 2025 United States
// Government as represented by the Administrator of NASA, Apache License 2.0,
// http://www.apache.org/licenses/LICENSE-2.0) because a drop-in has to; each such spot cites its
// ref: line.  The PointCloud2 message, the GPU objects and update() are new.  The reference text
// itself is not reproduced: integration/ carries the change as a patch against the reference tree.
#include "mujoco_cameras.hpp"

#include <algorithm>
#include <cmath>
#include <cstring>

namespace mujoco_ros2_control
{

namespace
{
// one sensor_msgs/Image header for a W x H image of `channel_bytes` bytes per pixel
void describe_image(sensor_msgs::msg::Image& img, const std::string& frame, uint32_t w, uint32_t h,
                    uint32_t bytes_per_pixel, const char* encoding)
{
  img.header.frame_id = frame;
  img.width = w;
  img.height = h;
  img.step = w * bytes_per_pixel;
  img.encoding = encoding;
  img.data.assign(static_cast<size_t>(w) * h * bytes_per_pixel, 0);
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

// One CameraData per mjModel camera (ref:src/mujoco_cameras.cpp:40-129): resolution and fovy from the
// model, topics and frame from the URDF <sensor> of the same name or the reference's defaults
// (<name>_frame, <name>/camera_info, <name>/color, <name>/depth, ref :71-74), pinhole intrinsics
// ref :119-124.  New: the organised PointCloud2 on <depth_topic>/points.
void MujocoCameras::register_cameras(const hardware_interface::HardwareInfo& hardware_info)
{
  cameras_.clear();
  for (int id = 0; id < mj_model_->ncam; ++id)
  {
    CameraData cam;
    cam.cam_id = id;
    cam.name = mj_model_->names + mj_model_->name_camadr[id];
    cam.width = static_cast<uint32_t>(mj_model_->cam_resolution[2 * id]);
    cam.height = static_cast<uint32_t>(mj_model_->cam_resolution[2 * id + 1]);

    const auto sensor = std::find_if(hardware_info.sensors.begin(), hardware_info.sensors.end(),
                                     [&](const hardware_interface::ComponentInfo& c) { return c.name == cam.name; });
    if (sensor != hardware_info.sensors.end())
    {
      // all four or an exception, as in the reference (.at(), ref :63-66)
      const auto& prm = sensor->parameters;
      cam.frame_name = prm.at("frame_name");
      cam.info_topic = prm.at("info_topic");
      cam.image_topic = prm.at("image_topic");
      cam.depth_topic = prm.at("depth_topic");
    }
    else
    {
      cam.frame_name = cam.name + "_frame";
      cam.info_topic = cam.name + "/camera_info";
      cam.image_topic = cam.name + "/color";
      cam.depth_topic = cam.name + "/depth";
    }
    cam.points_topic = cam.depth_topic + "/points";

    cam.camera_info_pub = node_->create_publisher<sensor_msgs::msg::CameraInfo>(cam.info_topic, 1);
    cam.image_pub = node_->create_publisher<sensor_msgs::msg::Image>(cam.image_topic, 1);
    cam.depth_image_pub = node_->create_publisher<sensor_msgs::msg::Image>(cam.depth_topic, 1);
    cam.cloud_pub = node_->create_publisher<sensor_msgs::msg::PointCloud2>(cam.points_topic, 1);

    const size_t npix = static_cast<size_t>(cam.width) * cam.height;
    describe_image(cam.image, cam.frame_name, cam.width, cam.height, 3, sensor_msgs::image_encodings::RGB8);
    describe_image(cam.depth_image, cam.frame_name, cam.width, cam.height, sizeof(float),
                   sensor_msgs::image_encodings::TYPE_32FC1);
    cam.image_buffer.assign(npix * 3, 0);
    cam.depth_buffer.assign(npix, 0.0f);

    // CameraInfo: plumb_bob with five zero coefficients, K and P from fovy (ref :119-124); R stays
    // all-zero exactly as the reference leaves it (ref :115)
    auto& info = cam.camera_info;
    info.header.frame_id = cam.frame_name;
    info.width = cam.width;
    info.height = cam.height;
    info.distortion_model = "plumb_bob";
    info.d.assign(5, 0.0);
    info.k.fill(0.0);
    info.r.fill(0.0);
    info.p.fill(0.0);
    const double f = (1.0 / std::tan((mj_model_->cam_fovy[id] * M_PI / 180.0) / 2.0)) * cam.height / 2.0;
    const double cx = static_cast<double>(cam.width) / 2.0, cy = static_cast<double>(cam.height) / 2.0;
    info.k[0] = info.k[4] = info.p[0] = info.p[5] = f;
    info.k[2] = info.p[2] = cx;
    info.k[5] = info.p[6] = cy;
    info.k[8] = info.p[10] = 1.0;

    // PointCloud2: organised, x/y/z FLOAT32 at 0/4/8, point_step 12, NaN for invalid depth
    auto& cloud = cam.cloud;
    cloud.header.frame_id = cam.frame_name;
    cloud.height = cam.height;
    cloud.width = cam.width;
    cloud.fields.resize(3);
    for (int k = 0; k < 3; ++k)
    {
      cloud.fields[k].name = std::string(1, "xyz"[k]);
      cloud.fields[k].offset = 4 * k;
      cloud.fields[k].datatype = sensor_msgs::msg::PointField::FLOAT32;
      cloud.fields[k].count = 1;
    }
    cloud.is_bigendian = false;
    cloud.is_dense = false;
    cloud.point_step = 12;
    cloud.row_step = cam.width * cloud.point_step;
    cloud.data.assign(npix * cloud.point_step, 0);

    RCLCPP_INFO(node_->get_logger(), "camera '%s': %ux%u fovy %.2f, frame %s, topics %s %s %s %s", cam.name.c_str(),
                cam.width, cam.height, static_cast<double>(mj_model_->cam_fovy[id]), cam.frame_name.c_str(),
                cam.image_topic.c_str(), cam.depth_topic.c_str(), cam.info_topic.c_str(), cam.points_topic.c_str());
    cameras_.push_back(std::move(cam));
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
      report_failure(std::string("b2_camera_create: ") + b2_last_error());
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

// One publish cycle (ref:src/mujoco_cameras.cpp:199-265).  With a GL readback hook the reference's
// steps 0-2 stay (render + mjr_readPixels) and step 3 -- float32 linearise + flip, ref :222-240 -- runs
// on the GPU together with the cloud projection.  Without OpenGL the depth buffer is ray-cast from
// the scene mirror: the camera frames are read and the render is ENQUEUED while the sim mutex is
// held (b2_camera_render_to_cloud_begin, nothing waits for the GPU), so frames and geom poses belong
// to the same physics step -- what mjv_copyData under the same mutex gives the reference (:205-208).
void MujocoCameras::update()
{
  const size_t nc = cameras_.size();
  std::vector<char> fresh(nc, 0), rgb_fresh(nc, 0);
  if (!gl_readback_)
  {
    pose_scratch_.resize(12 * nc);
    std::vector<char> begun(nc, 0);
    {
      std::unique_lock<std::recursive_mutex> lock(*sim_mutex_);
      if (refresh_in_update_ && mirror_ != nullptr)
      {
        mirror_->refresh(mj_data_);
      }
      for (size_t k = 0; k < nc; ++k)
      {
        auto& cam = cameras_[k];
        double* pos = &pose_scratch_[12 * k];
        std::copy_n(mj_data_->cam_xpos + 3 * cam.cam_id, 3, pos);
        std::copy_n(mj_data_->cam_xmat + 9 * cam.cam_id, 9, pos + 3);
        if (b2_camera_render_to_cloud_begin(cam.engine, 1, nullptr, pos, pos + 3, cam.depth_buffer.data(),
                                            reinterpret_cast<float*>(cam.depth_image.data.data()),
                                            cam.cloud.data.data()) == B2_OK)
        {
          begun[k] = 1;
        }
        else
        {
          report_failure(std::string("render_to_cloud_begin (") + cam.name + "): " + b2_last_error());
        }
      }
    }
    for (size_t k = 0; k < nc; ++k)
    {
      if (!begun[k])
      {
        continue;
      }
      if (b2_camera_render_to_cloud_end(cameras_[k].engine) == B2_OK)
      {
        fresh[k] = 1;
      }
      else
      {
        report_failure(std::string("render_to_cloud_end (") + cameras_[k].name + "): " + b2_last_error());
      }
    }
  }
  else
  {
    for (size_t k = 0; k < nc; ++k)
    {
      auto& cam = cameras_[k];
      if (!gl_readback_(cam))
      {
        report_failure("GL readback failed for camera " + cam.name);
        continue;
      }
      if (b2_depth_to_cloud(cam.engine, 1, cam.depth_buffer.data(), reinterpret_cast<float*>(cam.depth_image.data.data()),
                            cam.cloud.data.data()) != B2_OK)
      {
        report_failure(std::string("b2_depth_to_cloud (") + cam.name + "): " + b2_last_error());
        continue;
      }
      // RGB: GL's origin is bottom-left; the message wants the top row first (ref :243-249)
      const size_t row = static_cast<size_t>(cam.width) * 3;
      for (uint32_t h = 0; h < cam.height; ++h)
      {
        std::copy_n(&cam.image_buffer[h * row], row, &cam.image.data[(cam.height - 1 - h) * row]);
      }
      fresh[k] = rgb_fresh[k] = 1;
    }
  }

  // publish what this cycle produced: nothing stale goes out with a fresh stamp, and no colour image
  // exists unless a GL readback filled it (ray casting renders depth only)
  for (size_t k = 0; k < nc; ++k)
  {
    if (!fresh[k])
    {
      continue;
    }
    auto& cam = cameras_[k];
    const auto stamp = node_->now();
    cam.depth_image.header.stamp = stamp;
    cam.camera_info.header.stamp = stamp;
    cam.cloud.header.stamp = stamp;
    if (rgb_fresh[k])
    {
      cam.image.header.stamp = stamp;
      cam.image_pub->publish(cam.image);
    }
    cam.depth_image_pub->publish(cam.depth_image);
    cam.camera_info_pub->publish(cam.camera_info);
    cam.cloud_pub->publish(cam.cloud);
  }
}

void MujocoCameras::report_failure(const std::string& what)
{
  {
    std::lock_guard<std::mutex> g(error_mutex_);
    last_error_ = what;
  }
  if (failures_++ % 50 == 0)
  {
    RCLCPP_ERROR(node_->get_logger(), "%s", what.c_str());
  }
}

}  // namespace mujoco_ros2_control

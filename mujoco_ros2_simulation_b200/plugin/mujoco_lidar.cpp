// mujoco_lidar.cpp -- B200 drop-in for ref:src/mujoco_lidar.cpp (see mujoco_lidar.hpp).
#include "mujoco_lidar.hpp"

#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstring>

namespace mujoco_ros2_control
{

namespace
{
std::optional<hardware_interface::ComponentInfo>
get_sensor_from_info(const hardware_interface::HardwareInfo& hardware_info, const std::string& name)
{
  // ref:include/mujoco_ros2_control/utils.hpp:30-42
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

// "<lidar>-<idx>" -> (lidar, idx); ("", -1)-style failure exactly as ref:src/mujoco_lidar.cpp:30-49
std::pair<std::string, int> parse_lidar_name(const std::string& sensor_name)
{
  const auto split_idx = sensor_name.find_last_of("-");
  if (split_idx == std::string::npos)
  {
    return { sensor_name, -1 };
  }
  const auto lidar_name = sensor_name.substr(0, split_idx);
  const auto lidar_idx = sensor_name.substr(split_idx + 1);
  if (lidar_idx.empty() || !std::all_of(lidar_idx.begin(), lidar_idx.end(), ::isdigit))
  {
    return { lidar_name, -1 };
  }
  return { lidar_name, std::stoi(lidar_idx) };
}

// ref:src/mujoco_lidar.cpp:54-112: same required / optional parameters, defaults and message fields
std::optional<LidarData> get_lidar_data(const hardware_interface::HardwareInfo& hardware_info,
                                        const std::string& name)
{
  const auto sensor_info_maybe = get_sensor_from_info(hardware_info, name);
  if (!sensor_info_maybe.has_value())
  {
    return std::nullopt;
  }
  const auto sensor_info = sensor_info_maybe.value();
  auto get_parameter = [&](const std::string& key) -> std::optional<std::string> {
    if (auto it = sensor_info.parameters.find(key); it != sensor_info.parameters.end())
    {
      return it->second;
    }
    return std::nullopt;
  };

  auto frame_name = get_parameter("frame_name");
  auto min_angle = get_parameter("min_angle");
  auto max_angle = get_parameter("max_angle");
  auto angle_increment = get_parameter("angle_increment");
  auto laserscan_topic = get_parameter("laserscan_topic");
  auto range_min = get_parameter("range_min");
  auto range_max = get_parameter("range_max");
  if (!frame_name || !angle_increment || !min_angle || !max_angle)
  {
    return std::nullopt;
  }

  LidarData lidar_sensor;
  lidar_sensor.name = name;
  lidar_sensor.frame_name = *frame_name;
  lidar_sensor.min_angle = std::stod(*min_angle);
  lidar_sensor.max_angle = std::stod(*max_angle);
  lidar_sensor.angle_increment = std::stod(*angle_increment);
  // f64 truncation on purpose: 0.6 / 0.025 = 23.999.. -> 24 rays (ref :92-93)
  lidar_sensor.num_rangefinders =
      static_cast<int>((lidar_sensor.max_angle - lidar_sensor.min_angle) / lidar_sensor.angle_increment) + 1;
  lidar_sensor.laserscan_topic = laserscan_topic.has_value() ? laserscan_topic.value() : "/scan";
  lidar_sensor.range_min = range_min.has_value() ? std::stod(range_min.value()) : 0.0;
  lidar_sensor.range_max = range_max.has_value() ? std::stod(range_max.value()) : 1000.0;
  lidar_sensor.sensor_indexes.resize(lidar_sensor.num_rangefinders);
  lidar_sensor.local_dir.assign(3 * (size_t)lidar_sensor.num_rangefinders, 0.0);
  lidar_sensor.have_ray.assign((size_t)lidar_sensor.num_rangefinders, 0);

  lidar_sensor.laser_scan_msg.header.frame_id = lidar_sensor.frame_name;
  lidar_sensor.laser_scan_msg.time_increment = 0.0;
  lidar_sensor.laser_scan_msg.angle_min = lidar_sensor.min_angle;
  lidar_sensor.laser_scan_msg.angle_max = lidar_sensor.max_angle;  // configured max, not the last ray's
  lidar_sensor.laser_scan_msg.angle_increment = lidar_sensor.angle_increment;
  lidar_sensor.laser_scan_msg.range_min = lidar_sensor.range_min;
  lidar_sensor.laser_scan_msg.range_max = lidar_sensor.range_max;
  lidar_sensor.laser_scan_msg.ranges.resize(lidar_sensor.num_rangefinders);
  lidar_sensor.laser_scan_msg.intensities.resize(0);
  return lidar_sensor;
}

MujocoLidar::MujocoLidar(rclcpp::Node::SharedPtr& node, std::recursive_mutex* sim_mutex, mjData* mujoco_data,
                         mjModel* mujoco_model, double lidar_publish_rate, SceneMirror* mirror)
  : node_(node)
  , sim_mutex_(sim_mutex)
  , mj_data_(mujoco_data)
  , mj_model_(mujoco_model)
  , lidar_publish_rate_(lidar_publish_rate)
  , mirror_(mirror)
{
}

MujocoLidar::~MujocoLidar()
{
  close();
}

bool MujocoLidar::register_lidar(const hardware_interface::HardwareInfo& hardware_info)
{
  lidar_sensors_.resize(0);
  for (int i = 0; i < mj_model_->nsensor; ++i)
  {
    if (mj_model_->sensor_type[i] != mjtSensor::mjSENS_RANGEFINDER)
    {
      continue;
    }
    const auto sensor_name_maybe = mj_id2name(mj_model_, mjOBJ_SENSOR, i);
    if (sensor_name_maybe == nullptr)
    {
      continue;
    }
    const std::string sensor_name(sensor_name_maybe);
    const auto [lidar_name, idx] = parse_lidar_name(sensor_name);
    if (idx == -1)
    {
      continue;
    }
    auto lidar_it = std::find_if(lidar_sensors_.begin(), lidar_sensors_.end(),
                                 [&lidar_name](const LidarData& data) { return data.name == lidar_name; });
    if (lidar_it == lidar_sensors_.end())
    {
      auto new_data_maybe = get_lidar_data(hardware_info, lidar_name);
      if (!new_data_maybe.has_value())
      {
        last_error_ = "Failed to parse required configuration from ros2_control xacro: " + lidar_name;
        return false;
      }
      auto lidar = new_data_maybe.value();
      lidar.scan_pub = node_->create_publisher<sensor_msgs::msg::LaserScan>(lidar.laserscan_topic, 1);
      lidar.laser_scan_msg.scan_time = 1.0 / lidar_publish_rate_;
      lidar_sensors_.push_back(std::move(lidar));
      lidar_it = lidar_sensors_.end() - 1;
    }
    if (idx >= lidar_it->num_rangefinders)
    {
      continue;  // more replicas in the MJCF than the URDF parameters describe
    }
    lidar_it->sensor_indexes[idx] = mj_model_->sensor_adr[i];

    // GPU path: the ray of rangefinder i is the +Z axis of its site (engine_sensor.c), expressed
    // in the frame of the body carrying the site; that body is also the rangefinder's bodyexclude.
    const int site = mj_model_->sensor_objid[i];
    const int body = mj_model_->site_bodyid[site];
    if (lidar_it->bodyid < 0)
    {
      lidar_it->bodyid = body;
      std::memcpy(lidar_it->site_pos, mj_model_->site_pos + 3 * site, 3 * sizeof(double));
      lidar_it->cutoff = mj_model_->sensor_cutoff ? mj_model_->sensor_cutoff[i] : 0.0;
    }
    else if (lidar_it->bodyid != body)
    {
      last_error_ = "rangefinder sites of lidar '" + lidar_name + "' live on different bodies";
      return false;
    }
    const mjtNum* q = mj_model_->site_quat + 4 * site;
    const double n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    const double w = q[0] / n, x = q[1] / n, y = q[2] / n, z = q[3] / n;
    lidar_it->local_dir[3 * idx + 0] = 2 * (x * z + w * y);
    lidar_it->local_dir[3 * idx + 1] = 2 * (y * z - w * x);
    lidar_it->local_dir[3 * idx + 2] = 1 - 2 * (x * x + y * y);
    lidar_it->have_ray[idx] = 1;
  }
  return true;
}

bool MujocoLidar::create_engines()
{
  if (mirror_ == nullptr)
  {
    last_error_ = "MujocoLidar has no scene mirror: the B200 sensor path has no CPU fallback";
    return false;
  }
  for (auto& lidar : lidar_sensors_)
  {
    if (lidar.engine != nullptr)
    {
      continue;
    }
    b2_lidar_desc d{};
    d.nray = lidar.num_rangefinders;
    d.local_dir = lidar.local_dir.data();  // rays without a sensor stay zero-length: reported -1
    d.tile_w = 0;
    d.geomgroup = nullptr;                  // rangefinder: geomgroup NULL, flg_static 1
    d.flg_static = 1;
    d.bodyexclude = lidar.bodyid;
    d.cutoff = lidar.cutoff;
    d.range_min = lidar.range_min;
    d.range_max = lidar.range_max;
    if (b2_lidar_create(mirror_->handle(), &d, &lidar.engine) != B2_OK)
    {
      last_error_ = std::string("b2_lidar_create: ") + b2_last_error();
      return false;
    }
  }
  return true;
}

void MujocoLidar::init()
{
  if (!create_engines())
  {
    publish_lidar_ = false;
    return;
  }
  publish_lidar_ = true;
  rendering_thread_ = std::thread(&MujocoLidar::update_loop, this);
}

void MujocoLidar::close()
{
  publish_lidar_ = false;
  if (rendering_thread_.joinable())
  {
    rendering_thread_.join();
  }
  for (auto& lidar : lidar_sensors_)
  {
    if (lidar.engine != nullptr)
    {
      b2_lidar_destroy(lidar.engine);
      lidar.engine = nullptr;
    }
  }
}

void MujocoLidar::update_loop()
{
  rclcpp::Rate rate(lidar_publish_rate_);
  while (rclcpp::ok() && publish_lidar_)
  {
    update();
    rate.sleep();
  }
}

void MujocoLidar::update()
{
  // Step 1: lock the sim and take what the cast needs: the pose of each lidar's body (24 doubles
  // per lidar instead of the whole sensordata array) and, unless the physics thread already did it
  // after mj_step, the geom poses for the scene mirror.
  std::vector<double> pos(3 * lidar_sensors_.size()), mat(9 * lidar_sensors_.size());
  {
    std::unique_lock<std::recursive_mutex> lock(*sim_mutex_);
    if (refresh_in_update_)
    {
      mirror_->refresh(mj_data_);
    }
    for (size_t k = 0; k < lidar_sensors_.size(); ++k)
    {
      const auto& lidar = lidar_sensors_[k];
      const mjtNum* R = mj_data_->xmat + 9 * lidar.bodyid;
      const mjtNum* p = mj_data_->xpos + 3 * lidar.bodyid;
      for (int r = 0; r < 3; ++r)
      {
        pos[3 * k + r] = p[r] + R[3 * r] * lidar.site_pos[0] + R[3 * r + 1] * lidar.site_pos[1] +
                         R[3 * r + 2] * lidar.site_pos[2];
      }
      std::memcpy(&mat[9 * k], R, 9 * sizeof(double));
    }
  }

  // Step 2: cast.  The kernel's epilogue applies the rangefinder cutoff, the double -> float cast
  // and the range filter, so ranges land in the message as the reference would have filled them.
  for (size_t k = 0; k < lidar_sensors_.size(); ++k)
  {
    auto& lidar = lidar_sensors_[k];
    if (b2_lidar_cast(lidar.engine, 1, nullptr, &pos[3 * k], &mat[9 * k], lidar.laser_scan_msg.ranges.data(),
                      nullptr, nullptr, nullptr) != B2_OK)
    {
      last_error_ = std::string("b2_lidar_cast: ") + b2_last_error();
      return;
    }
  }

  // Step 3: publish
  for (auto& lidar : lidar_sensors_)
  {
    lidar.laser_scan_msg.header.stamp = node_->now();
    lidar.scan_pub->publish(lidar.laser_scan_msg);
  }
}

}  // namespace mujoco_ros2_control

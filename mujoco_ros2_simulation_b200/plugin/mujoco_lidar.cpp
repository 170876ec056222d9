// mujoco_lidar.cpp -- B200 drop-in for ref:src/mujoco_lidar.cpp (see mujoco_lidar.hpp).
//
// Interface-compatible re-implementation.  The class name, its public methods, the URDF sensor
// parameter names and defaults, the "<lidar>-<idx>" sensor naming and the LaserScan field values
// follow the reference file (yassine-cherni/mujoco_ros2_simulation, src/mujoco_lidar.cpp, Copyright
// (c) 2025 United States Government as represented by the Administrator of NASA, Apache License 2.0,
// http://www.apache.org/licenses/LICENSE-2.0) because a drop-in has to; each such spot cites its
// ref: line.  The ray tables, the GPU objects and update() are new.  The reference text itself is
// not reproduced: integration/ carries the change as a patch against the reference tree.
#include "mujoco_lidar.hpp"

#include <algorithm>
#include <cmath>
#include <cstring>

namespace mujoco_ros2_control
{

namespace
{
// ref:include/mujoco_ros2_control/utils.hpp:30-42 (get_sensor_from_info): the <sensor> block of the
// ros2_control tag with this name, or nullptr
const hardware_interface::ComponentInfo* find_sensor(const hardware_interface::HardwareInfo& info,
                                                     const std::string& name)
{
  const auto it = std::find_if(info.sensors.begin(), info.sensors.end(),
                               [&](const hardware_interface::ComponentInfo& c) { return c.name == name; });
  return it == info.sensors.end() ? nullptr : &*it;
}

// a sensor <param>: text, or nullptr when absent
const std::string* param(const hardware_interface::ComponentInfo& sensor, const char* key)
{
  const auto it = sensor.parameters.find(key);
  return it == sensor.parameters.end() ? nullptr : &it->second;
}

double param_or(const hardware_interface::ComponentInfo& sensor, const char* key, double fallback)
{
  const std::string* v = param(sensor, key);
  return v ? std::stod(*v) : fallback;
}
}  // namespace

// "<lidar>-<idx>": split at the LAST '-', the suffix must be all digits.  Failure is reported as
// idx = -1 (ref:src/mujoco_lidar.cpp:30-49; names come from <replicate sep="-">,
// ref:scripts/make_mjcf_from_robot_description.py:736-750).
std::pair<std::string, int> parse_lidar_name(const std::string& sensor_name)
{
  const size_t dash = sensor_name.rfind('-');
  if (dash == std::string::npos)
  {
    return { sensor_name, -1 };
  }
  const std::string head = sensor_name.substr(0, dash);
  int idx = 0;
  size_t ndigit = 0;
  for (size_t i = dash + 1; i < sensor_name.size(); ++i, ++ndigit)
  {
    const char ch = sensor_name[i];
    if (ch < '0' || ch > '9')
    {
      return { head, -1 };
    }
    idx = idx * 10 + (ch - '0');
  }
  return { head, ndigit == 0 ? -1 : idx };
}

// The static part of one lidar from its URDF <sensor> parameters.  Required: frame_name, min_angle,
// max_angle, angle_increment (ref :80); optional with the reference's defaults: laserscan_topic
// "/scan", range_min 0.0, range_max 1000.0 (ref :95-97).  LaserScan constants as ref :101-109.
std::optional<LidarData> get_lidar_data(const hardware_interface::HardwareInfo& hardware_info,
                                        const std::string& name)
{
  const hardware_interface::ComponentInfo* sensor = find_sensor(hardware_info, name);
  if (sensor == nullptr)
  {
    return std::nullopt;
  }
  for (const char* required : { "frame_name", "min_angle", "max_angle", "angle_increment" })
  {
    if (param(*sensor, required) == nullptr)
    {
      return std::nullopt;
    }
  }
  LidarData out;
  out.name = name;
  out.frame_name = *param(*sensor, "frame_name");
  out.min_angle = param_or(*sensor, "min_angle", 0.0);
  out.max_angle = param_or(*sensor, "max_angle", 0.0);
  out.angle_increment = param_or(*sensor, "angle_increment", 1.0);
  out.range_min = param_or(*sensor, "range_min", 0.0);
  out.range_max = param_or(*sensor, "range_max", 1000.0);
  const std::string* topic = param(*sensor, "laserscan_topic");
  out.laserscan_topic = topic ? *topic : "/scan";
  // the count relies on f64 truncation, as the reference's does: 0.6 / 0.025 = 23.999.. -> 24 rays
  // (ref :92-93)
  out.num_rangefinders = static_cast<int>((out.max_angle - out.min_angle) / out.angle_increment) + 1;
  const size_t n = static_cast<size_t>(out.num_rangefinders);
  out.sensor_indexes.assign(n, 0);
  out.local_dir.assign(3 * n, 0.0);
  out.have_ray.assign(n, 0);
  out.ray_cutoff.assign(n, 0.0);

  auto& msg = out.laser_scan_msg;
  msg.header.frame_id = out.frame_name;
  msg.angle_min = out.min_angle;
  msg.angle_max = out.max_angle;  // the CONFIGURED maximum, which may lie past the last ray (ref :104)
  msg.angle_increment = out.angle_increment;
  msg.time_increment = 0.0;
  msg.range_min = out.range_min;
  msg.range_max = out.range_max;
  msg.ranges.assign(n, 0.0f);
  msg.intensities.clear();
  return out;
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

// Groups the model's rangefinder sensors "<lidar>-<idx>" into lidars (ref :124-210) and, new here,
// records for every ray what the GPU cast needs: its direction in the frame of the body that carries
// the rangefinder sites (the site's +Z axis, engine_sensor.c), that body (the rangefinder's
// bodyexclude) and the sensor cutoff.
bool MujocoLidar::register_lidar(const hardware_interface::HardwareInfo& hardware_info)
{
  lidar_sensors_.clear();
  const auto lidar_named = [this](const std::string& n) {
    return std::find_if(lidar_sensors_.begin(), lidar_sensors_.end(),
                        [&n](const LidarData& d) { return d.name == n; });
  };
  for (int i = 0; i < mj_model_->nsensor; ++i)
  {
    if (mj_model_->sensor_type[i] != mjtSensor::mjSENS_RANGEFINDER)
    {
      continue;
    }
    const char* raw_name = mj_id2name(mj_model_, mjOBJ_SENSOR, i);
    if (raw_name == nullptr)
    {
      RCLCPP_WARN(node_->get_logger(), "rangefinder sensor %d has no name: not part of any lidar", i);
      continue;
    }
    const auto [lidar_name, idx] = parse_lidar_name(raw_name);
    if (idx < 0)
    {
      RCLCPP_WARN(node_->get_logger(), "rangefinder '%s' is not of the form <lidar>-<idx>: skipped", raw_name);
      continue;
    }
    auto lidar = lidar_named(lidar_name);
    if (lidar == lidar_sensors_.end())
    {
      auto parsed = get_lidar_data(hardware_info, lidar_name);
      if (!parsed)
      {
        last_error_ = "Failed to parse required configuration from ros2_control xacro: " + lidar_name;
        RCLCPP_ERROR(node_->get_logger(), "%s", last_error_.c_str());
        return false;
      }
      parsed->scan_pub = node_->create_publisher<sensor_msgs::msg::LaserScan>(parsed->laserscan_topic, 1);
      parsed->laser_scan_msg.scan_time = 1.0 / lidar_publish_rate_;  // ref :175
      RCLCPP_INFO(node_->get_logger(), "lidar '%s': frame %s, %d rays from %.4f step %.4f rad, topic %s, range [%.3f, %.3f]",
                  parsed->name.c_str(), parsed->frame_name.c_str(), parsed->num_rangefinders, parsed->min_angle,
                  parsed->angle_increment, parsed->laserscan_topic.c_str(), parsed->range_min, parsed->range_max);
      lidar_sensors_.push_back(std::move(*parsed));
      lidar = lidar_sensors_.end() - 1;
    }
    if (idx >= lidar->num_rangefinders)
    {
      RCLCPP_WARN(node_->get_logger(), "lidar '%s': replica %d lies beyond the %d rays its URDF parameters describe: dropped",
                  lidar_name.c_str(), idx, lidar->num_rangefinders);
      continue;
    }
    lidar->sensor_indexes[idx] = mj_model_->sensor_adr[i];

    const int site = mj_model_->sensor_objid[i];
    const int body = mj_model_->site_bodyid[site];
    if (lidar->bodyid < 0)
    {
      lidar->bodyid = body;
      std::copy_n(mj_model_->site_pos + 3 * site, 3, lidar->site_pos);
    }
    else if (lidar->bodyid != body)
    {
      last_error_ = "rangefinder sites of lidar '" + lidar_name + "' live on different bodies";
      RCLCPP_ERROR(node_->get_logger(), "%s", last_error_.c_str());
      return false;
    }
    // third column of the site's rotation in the body frame, from its (normalised) quaternion
    const mjtNum* q = mj_model_->site_quat + 4 * site;
    const double inv = 1.0 / std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    const double w = q[0] * inv, x = q[1] * inv, y = q[2] * inv, z = q[3] * inv;
    double* dir = &lidar->local_dir[3 * idx];
    dir[0] = 2 * (x * z + w * y);
    dir[1] = 2 * (y * z - w * x);
    dir[2] = 1 - 2 * (x * x + y * y);
    lidar->have_ray[idx] = 1;
    lidar->ray_cutoff[idx] = mj_model_->sensor_cutoff ? mj_model_->sensor_cutoff[i] : 0.0;
  }
  // MuJoCo applies sensor_cutoff per sensor; the cast takes one value per lidar (every replica of a
  // <replicate>d rangefinder shares it).  Refuse a model where that does not hold.
  for (auto& lidar : lidar_sensors_)
  {
    bool first = true;
    for (int k = 0; k < lidar.num_rangefinders; ++k)
    {
      if (!lidar.have_ray[k])
      {
        continue;
      }
      if (first)
      {
        lidar.cutoff = lidar.ray_cutoff[k];
        first = false;
      }
      else if (lidar.ray_cutoff[k] != lidar.cutoff)
      {
        last_error_ = "rangefinders of lidar '" + lidar.name + "' have different cutoffs";
        RCLCPP_ERROR(node_->get_logger(), "%s", last_error_.c_str());
        return false;
      }
    }
    const auto missing = std::count(lidar.have_ray.begin(), lidar.have_ray.end(), 0);
    if (missing > 0)
    {
      RCLCPP_WARN(node_->get_logger(), "lidar '%s': %d of %d rays have no rangefinder sensor in the model (reported as -1)",
                  lidar.name.c_str(), static_cast<int>(missing), lidar.num_rangefinders);
    }
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

// One publish cycle.  ref:src/mujoco_lidar.cpp:244-280 locks the sim, copies sensordata, gathers and
// filters; here the sim is locked for the same short moment: the sensor frames are read and every
// lidar's cast is ENQUEUED (b2_lidar_cast_begin: pose refresh, cast, read-back -- nothing waits for
// the GPU) while the lock is held, so the frames and the geom poses the rays meet belong to the same
// physics step, as they do when MuJoCo computes the rangefinders inside mj_step.  Waiting and
// publishing happen after the unlock.
void MujocoLidar::update()
{
  const size_t nl = lidar_sensors_.size();
  pose_scratch_.resize(12 * nl);
  std::vector<char> begun(nl, 0);
  {
    std::unique_lock<std::recursive_mutex> lock(*sim_mutex_);
    if (refresh_in_update_)
    {
      mirror_->refresh(mj_data_);
    }
    for (size_t k = 0; k < nl; ++k)
    {
      auto& lidar = lidar_sensors_[k];
      const mjtNum* R = mj_data_->xmat + 9 * lidar.bodyid;
      const mjtNum* p = mj_data_->xpos + 3 * lidar.bodyid;
      double* pos = &pose_scratch_[12 * k];
      double* mat = pos + 3;
      for (int r = 0; r < 3; ++r)
      {
        pos[r] = p[r] + R[3 * r] * lidar.site_pos[0] + R[3 * r + 1] * lidar.site_pos[1] +
                 R[3 * r + 2] * lidar.site_pos[2];
      }
      std::copy_n(R, 9, mat);
      // the kernel's epilogue applies the rangefinder cutoff, the double -> float cast and the range
      // filter (ref :262, :269-271): ranges land in the message as the reference fills them
      if (b2_lidar_cast_begin(lidar.engine, 1, nullptr, pos, mat, lidar.laser_scan_msg.ranges.data(), nullptr,
                              nullptr, nullptr) == B2_OK)
      {
        begun[k] = 1;
      }
      else
      {
        report_failure(std::string("b2_lidar_cast_begin (") + lidar.name + "): " + b2_last_error());
      }
    }
  }
  for (size_t k = 0; k < nl; ++k)
  {
    auto& lidar = lidar_sensors_[k];
    if (!begun[k])
    {
      continue;
    }
    if (b2_lidar_cast_end(lidar.engine) != B2_OK)
    {
      report_failure(std::string("b2_lidar_cast_end (") + lidar.name + "): " + b2_last_error());
      continue;  // nothing stale is published for this lidar
    }
    lidar.laser_scan_msg.header.stamp = node_->now();
    lidar.scan_pub->publish(lidar.laser_scan_msg);
  }
}

void MujocoLidar::report_failure(const std::string& what)
{
  {
    std::lock_guard<std::mutex> g(error_mutex_);
    last_error_ = what;
  }
  // throttled: a persistent failure in a 5 Hz loop must not flood the log
  if (failures_++ % 50 == 0)
  {
    RCLCPP_ERROR(node_->get_logger(), "%s", what.c_str());
  }
}

}  // namespace mujoco_ros2_control

// scene_mirror.hpp -- the GPU scene mirror as the plugin sees it (new; no reference counterpart).
//
// MujocoSystemInterface::on_init builds one SceneMirror right after LoadModel
// (ref:src/mujoco_system_interface.cpp:677) and hands it to MujocoLidar / MujocoCameras next to the
// mjModel* / mjData* / mutex* they already receive (:707-717).  The physics thread calls refresh()
// right after mj_step (:1691, :1731) while it still holds sim_mutex_: it only copies geom_xpos /
// geom_xmat into pinned staging and enqueues the upload + BVH refit (b2_scene_update_poses).
#ifndef B2_SCENE_MIRROR_HPP
#define B2_SCENE_MIRROR_HPP

#include <stdexcept>
#include <string>

#include "../../include/b2sense.h"
#include "b2_compat.hpp"

namespace mujoco_ros2_control
{

class SceneMirror
{
public:
  SceneMirror(const mjModel* m, int device = 0) : model_(m)
  {
    b2_model_desc d{};
    d.ngeom = m->ngeom;
    d.nbody = m->nbody;
    d.nmesh = m->nmesh;
    d.nmat = m->nmat;
    d.geom_type = m->geom_type;
    d.geom_bodyid = m->geom_bodyid;
    d.geom_group = m->geom_group;
    d.geom_matid = m->geom_matid;
    d.geom_dataid = m->geom_dataid;
    d.geom_size = m->geom_size;
    d.geom_rgba = m->geom_rgba;
    d.mat_rgba = m->mat_rgba;
    d.body_weldid = m->body_weldid;
    d.mesh_vertadr = m->mesh_vertadr;
    d.mesh_vertnum = m->mesh_vertnum;
    d.mesh_faceadr = m->mesh_faceadr;
    d.mesh_facenum = m->mesh_facenum;
    d.mesh_vert = m->mesh_vert;
    d.mesh_face = m->mesh_face;
    if (b2_scene_create(&d, 1, device, nullptr, &scene_) != B2_OK)
    {
      throw std::runtime_error(std::string("b2_scene_create: ") + b2_last_error());
    }
  }
  ~SceneMirror()
  {
    b2_scene_destroy(scene_);
  }
  SceneMirror(const SceneMirror&) = delete;
  SceneMirror& operator=(const SceneMirror&) = delete;

  // physics thread, under sim_mutex_, after mj_step
  bool refresh(const mjData* d)
  {
    return b2_scene_update_poses(scene_, 0, 1, d->geom_xpos, d->geom_xmat) == B2_OK;
  }

  b2_scene* handle() const
  {
    return scene_;
  }
  const mjModel* model() const
  {
    return model_;
  }

private:
  const mjModel* model_;
  b2_scene* scene_{ nullptr };
};

}  // namespace mujoco_ros2_control
#endif

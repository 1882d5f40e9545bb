//! src/flatten.rs (new file) — the scene graph as the flat tables of `rc_scene` (include/racer_cuda.h).
//!
//! `RenderData.scene` is a `&dyn Hittable` (src/renderer.rs:92-99): a tree of boxed trait objects that can be
//! traced but not enumerated.  The CUDA backend needs structure-of-arrays buffers, so every node of that tree
//! learns to describe itself: `Hittable::flatten` (the BVH and its nodes, src/bvh_node.rs), `SceneObject` and
//! `HittableSceneObject::flatten` (the eight geometries, src/geometry/*.rs), `Material::flatten`
//! (src/material/*.rs), `Texture::flatten` (src/texture/*.rs), `BackgroundColor::flatten`
//! (src/background_color.rs).  This file holds what they write into.
//!
//! Conventions of the C ABI that the implementations follow:
//!   * primitives are appended in the BVH's depth-first leaf order, so an `rc_bvh_node` leaf is a contiguous run;
//!   * `prim_id = (SceneObject::id() of the TOP-LEVEL object << 3) | side` (side = 0..5 for the faces of a
//!     `Boxx`, src/geometry/box.rs:22-71; 0 otherwise); 0 is reserved for "no hit" (src/renderer.rs:78-88);
//!   * `prim_aabb` is the stored `Aabb` of the top-level object — the volume `Node::hit` tests before it
//!     descends (src/bvh_node.rs:119), which for `RotateY` is NOT a bounding box of the rotated object;
//!   * wrappers: `Translate(RotateY(object))`, `Translate(object)` and `RotateY(object)` are expressible as one
//!     `rc_instance`; any other nesting is refused with an error (no silent approximation).
use std::collections::HashMap;
use std::sync::Arc;

use racer_cuda_sys as sys;

use crate::{aabb::Aabb, error::TracerError, material::Material, texture::Texture, vec3::Vec3};

pub fn v3(v: &Vec3) -> [f64; 3] {
    [*v.x(), *v.y(), *v.z()]
}

/// The wrappers met on the way down to a primitive, outermost first.
#[derive(Clone, Copy, Default, PartialEq)]
pub struct Wrap {
    /// (sin, cos) of `RotateY` (src/geometry/rotate_y.rs:19-27)
    pub rotate: Option<(f64, f64)>,
    /// offset of `Translate` (src/geometry/translate.rs:17-21)
    pub translate: Option<[f64; 3]>,
}

impl Wrap {
    pub fn is_none(&self) -> bool {
        self.rotate.is_none() && self.translate.is_none()
    }
}

/// The top-level scene object a primitive belongs to.
pub struct Top<'a> {
    pub id: usize,
    pub aabb: &'a Aabb,
}

#[derive(Default)]
pub struct FlatScene {
    pub prim_type: Vec<i32>,
    pub prim_data: Vec<f64>,     // 5 per primitive
    pub prim_material: Vec<i32>,
    pub prim_id: Vec<u32>,
    pub prim_instance: Vec<i32>, // -1 or index into `instances`
    pub prim_aabb: Vec<f64>,     // 6 per primitive
    pub prim_motion: Vec<f64>,   // 5 per primitive: pos_b, time_a, time_b (src/geometry/moving_sphere.rs:19-24)
    pub any_motion: bool,
    pub instances: Vec<sys::rc_instance>,
    wraps: Vec<Wrap>,
    pub materials: Vec<sys::rc_material>,
    pub textures: Vec<sys::rc_texture>,
    pub images: Vec<(i32, i32, Vec<u8>)>,
    pub perlin: Vec<sys::rc_perlin>,
    pub nodes: Vec<sys::rc_bvh_node>,
    pub bg_type: i32,
    pub bg_a: [f64; 3],
    pub bg_b: [f64; 3],
    material_index: HashMap<usize, i32>, // Arc data pointer -> index: shared materials are emitted once
    texture_index: HashMap<usize, i32>,
}

pub fn unsupported(what: &str) -> TracerError {
    TracerError::CudaBackend(-1, format!("scene not expressible for the CUDA backend: {}", what))
}

impl FlatScene {
    pub fn n_prims(&self) -> usize {
        self.prim_type.len()
    }

    fn instance(&mut self, wrap: &Wrap) -> i32 {
        if wrap.is_none() {
            return -1;
        }
        if let Some(i) = self.wraps.iter().position(|w| w == wrap) {
            return i as i32;
        }
        let (sin_theta, cos_theta) = wrap.rotate.unwrap_or((0.0, 1.0));
        self.instances.push(sys::rc_instance {
            flags: (wrap.rotate.is_some() as i32) | ((wrap.translate.is_some() as i32) << 1),
            reserved: 0,
            sin_theta,
            cos_theta,
            offset: wrap.translate.unwrap_or([0.0; 3]),
        });
        self.wraps.push(*wrap);
        (self.instances.len() - 1) as i32
    }

    #[allow(clippy::too_many_arguments)]
    pub fn push_prim(
        &mut self,
        top: &Top,
        side: u32,
        wrap: &Wrap,
        prim_type: i32,
        data: [f64; 5],
        motion: Option<[f64; 5]>,
        material: &Arc<dyn Material>,
    ) -> Result<(), TracerError> {
        if top.id == 0 || top.id >= (1 << 29) || side > 7 {
            return Err(unsupported("object id out of range"));
        }
        let material = self.material(material)?;
        let instance = self.instance(wrap);
        self.prim_type.push(prim_type);
        self.prim_data.extend_from_slice(&data);
        self.prim_material.push(material);
        self.prim_id.push(((top.id as u32) << 3) | side);
        self.prim_instance.push(instance);
        self.prim_aabb.extend_from_slice(&v3(top.aabb.min()));
        self.prim_aabb.extend_from_slice(&v3(top.aabb.max()));
        self.any_motion |= motion.is_some();
        self.prim_motion.extend_from_slice(&motion.unwrap_or([0.0; 5]));
        Ok(())
    }

    pub fn material(&mut self, m: &Arc<dyn Material>) -> Result<i32, TracerError> {
        let key = Arc::as_ptr(m) as *const () as usize;
        if let Some(i) = self.material_index.get(&key) {
            return Ok(*i);
        }
        let flat = m.flatten(self)?;
        self.materials.push(flat);
        let i = (self.materials.len() - 1) as i32;
        self.material_index.insert(key, i);
        Ok(i)
    }

    pub fn texture(&mut self, t: &Arc<dyn Texture>) -> Result<i32, TracerError> {
        let key = Arc::as_ptr(t) as *const () as usize;
        if let Some(i) = self.texture_index.get(&key) {
            return Ok(*i);
        }
        let flat = t.flatten(self)?; // (a Checkered texture pushes its two children first)
        self.textures.push(flat);
        let i = (self.textures.len() - 1) as i32;
        self.texture_index.insert(key, i);
        Ok(i)
    }

    /// The borrowed view `rc_upload_scene` takes; `images` must outlive the call.
    pub fn as_rc_scene(&self, images: &[sys::rc_image]) -> sys::rc_scene {
        sys::rc_scene {
            n_prims: self.prim_type.len() as i32,
            prim_type: self.prim_type.as_ptr(),
            prim_data: self.prim_data.as_ptr(),
            prim_material: self.prim_material.as_ptr(),
            prim_id: self.prim_id.as_ptr(),
            prim_instance: self.prim_instance.as_ptr(),
            prim_aabb: self.prim_aabb.as_ptr(),
            n_instances: self.instances.len() as i32,
            instances: self.instances.as_ptr(),
            n_materials: self.materials.len() as i32,
            materials: self.materials.as_ptr(),
            n_textures: self.textures.len() as i32,
            textures: self.textures.as_ptr(),
            n_images: images.len() as i32,
            images: images.as_ptr(),
            n_perlin: self.perlin.len() as i32,
            perlin: self.perlin.as_ptr(),
            n_nodes: self.nodes.len() as i32,
            nodes: self.nodes.as_ptr(),
            bg_type: self.bg_type,
            reserved: 0,
            bg_a: self.bg_a,
            bg_b: self.bg_b,
            prim_motion: if self.any_motion { self.prim_motion.as_ptr() } else { std::ptr::null() },
        }
    }
}

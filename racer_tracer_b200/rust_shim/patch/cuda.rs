//! src/renderer/cuda.rs — `impl Renderer for CudaRenderer`, the drop-in behind
//! `Renderer::render` (src/renderer.rs:101-107).  Source only: not compiled in this repository.
//!
//! The one invasive host change: `RenderData.scene` is `&dyn Hittable`, which cannot be
//! enumerated, so the shim keeps the `Vec<SceneObject>` the BVH was built from and asks every
//! object / material / texture / background to describe itself (`Flatten` below) into the SoA
//! buffers of `rc_scene`.  The scene is re-flattened and re-uploaded when `bvh.changed()`
//! (src/main.rs:178-183).
use std::sync::Mutex;

use racer_cuda_sys as sys;
use synchronoise::SignalEvent;

use crate::{
    config::RenderConfig, data_bus::DataWriter, error::TracerError, image_buffer::ImageBufferEvent,
    renderer::{RenderData, Renderer}, vec3::Vec3,
};

/// What a scene object contributes to the flat scene (implemented for Sphere, XyRect, XzRect,
/// YzRect, Boxx, RotateY, Translate; materials and textures have the analogous `describe`).
pub trait Flatten {
    fn flatten(&self, out: &mut FlatScene);
}

#[derive(Default)]
pub struct FlatScene {
    pub prim_type: Vec<i32>, pub prim_data: Vec<f64>, pub prim_material: Vec<i32>, pub prim_id: Vec<u32>,
    pub prim_instance: Vec<i32>, pub prim_aabb: Vec<f64>,
    pub prim_motion: Vec<f64>,   // 5 per prim (pos_b, time_a, time_b): MovingSphere, src/geometry/moving_sphere.rs
    pub instances: Vec<sys::rc_instance>,
    pub materials: Vec<sys::rc_material>, pub textures: Vec<sys::rc_texture>,
    pub images: Vec<(u32, u32, Vec<u8>)>, pub perlin: Vec<sys::rc_perlin>, pub nodes: Vec<sys::rc_bvh_node>,
    pub bg_type: i32, pub bg_a: [f64; 3], pub bg_b: [f64; 3],
}

pub struct CudaRenderer {
    config: RenderConfig,
    ctx: Mutex<*mut sys::rc_ctx>,   // used by one thread at a time (src/main.rs:163-199)
    seed: u64,
}
unsafe impl Send for CudaRenderer {}
unsafe impl Sync for CudaRenderer {}

fn check(status: i32) -> Result<(), TracerError> {
    if status == sys::RC_OK { return Ok(()); }
    if status == sys::RC_ERR_CANCELLED { return Err(TracerError::CancelEvent); }
    let msg = unsafe { std::ffi::CStr::from_ptr(sys::rc_last_error()) }.to_string_lossy().into_owned();
    Err(TracerError::CudaBackend(status, msg))   // new variant, exit code 23 (src/error.rs:71-97)
}

impl CudaRenderer {
    pub fn new(config: RenderConfig, devices: &[i32], seed: u64) -> Result<Self, TracerError> {
        let mut ctx = std::ptr::null_mut();
        check(unsafe { sys::rc_create(devices.as_ptr(), devices.len() as i32, &mut ctx) })?;
        Ok(Self { config, ctx: Mutex::new(ctx), seed })
    }

    /// Called from the render task when `bvh.changed()`: flatten + upload.
    pub fn upload(&self, flat: &FlatScene) -> Result<(), TracerError> {
        let images: Vec<sys::rc_image> = flat.images.iter()
            .map(|(w, h, px)| sys::rc_image { width: *w as i32, height: *h as i32, rgba: px.as_ptr() }).collect();
        let scene = sys::rc_scene {
            n_prims: flat.prim_type.len() as i32,
            prim_type: flat.prim_type.as_ptr(), prim_data: flat.prim_data.as_ptr(),
            prim_material: flat.prim_material.as_ptr(), prim_id: flat.prim_id.as_ptr(),
            prim_instance: flat.prim_instance.as_ptr(), prim_aabb: flat.prim_aabb.as_ptr(),
            n_instances: flat.instances.len() as i32, instances: flat.instances.as_ptr(),
            n_materials: flat.materials.len() as i32, materials: flat.materials.as_ptr(),
            n_textures: flat.textures.len() as i32, textures: flat.textures.as_ptr(),
            n_images: images.len() as i32, images: images.as_ptr(),
            n_perlin: flat.perlin.len() as i32, perlin: flat.perlin.as_ptr(),
            n_nodes: flat.nodes.len() as i32, nodes: flat.nodes.as_ptr(),
            bg_type: flat.bg_type, reserved: 0, bg_a: flat.bg_a, bg_b: flat.bg_b,
            prim_motion: if flat.prim_motion.is_empty() { std::ptr::null() } else { flat.prim_motion.as_ptr() },
        };
        check(unsafe { sys::rc_upload_scene(*self.ctx.lock().unwrap(), &scene) })
    }
}

impl Renderer for CudaRenderer {
    fn render(&self, rd: RenderData, writer: &DataWriter<ImageBufferEvent>) -> Result<(), TracerError> {
        let ctx = *self.ctx.lock().unwrap();
        let cam = rd.camera_data.to_rc_camera();           // the 14 fields, src/camera.rs:57-72
        check(unsafe { sys::rc_set_camera(ctx, &cam) })?;
        let (w, h) = (rd.image.width, rd.image.height);
        let params = sys::rc_params {
            width: w as i32, height: h as i32,
            samples: rd.config.render.samples as i32,       // src/renderer/cpu.rs:38
            max_depth: rd.config.render.max_depth as i32,   // :46
            seed: self.seed, world: 1, ..Default::default()
        };
        // SignalEvent -> int flag: a tiny watcher thread sets it (synchronoise has no raw handle)
        let cancel = CancelFlag::watch(rd.cancel_event);
        let mut rgb = vec![0f64; w * h * 3];
        check(unsafe { sys::rc_render(ctx, &params, rgb.as_mut_ptr(), cancel.as_ptr()) })?;
        if cancel.is_set() { return Ok(()); }               // cancelled renders write nothing, cpu.rs:55-62
        // few, large BufferUpdate messages: Bus::new(1024) + try_broadcast is fatal when full (H6).
        // One message per band of rows; `self.config.samples` already normalised the sum on the device.
        let band = (h + 15) / 16;
        for r in (0..h).step_by(band) {
            let rows = band.min(h - r);
            let px: Vec<Vec3> = rgb[r * w * 3..(r + rows) * w * 3].chunks_exact(3)
                .map(|c| Vec3::new(c[0], c[1], c[2])).collect();
            writer.write(ImageBufferEvent::BufferUpdate { rgb: px, r, c: 0, width: w, height: rows })?;
        }
        Ok(())
    }
}

/// `preview_renderer: CudaPreview` — the drop-in for CpuRendererScaled (src/renderer/cpu_scaled.rs).
pub struct CudaPreviewRenderer { inner: CudaRenderer, scale_width: usize, scale_height: usize }

fn get_highest_divdable(value: usize, mut div: usize) -> usize {   // cpu_scaled.rs:17-23
    while (value % div) != 0 { div -= 1; }
    div
}

impl CudaPreviewRenderer {
    pub fn new(config: RenderConfig, image: &Image, devices: &[i32], seed: u64) -> Result<Self, TracerError> {
        let scale_width = get_highest_divdable(image.width / config.num_threads_width, config.scale);     // cpu_scaled.rs:31-34
        let scale_height = get_highest_divdable(image.height / config.num_threads_height, config.scale);
        Ok(Self { inner: CudaRenderer::new(config, devices, seed)?, scale_width, scale_height })
    }
}

impl Renderer for CudaPreviewRenderer {
    fn render(&self, rd: RenderData, writer: &DataWriter<ImageBufferEvent>) -> Result<(), TracerError> {
        let ctx = *self.inner.ctx.lock().unwrap();
        check(unsafe { sys::rc_set_camera(ctx, &rd.camera_data.to_rc_camera()) })?;
        let (w, h) = (rd.image.width, rd.image.height);
        let params = sys::rc_params {
            width: w as i32, height: h as i32,
            samples: self.inner.config.samples as i32, max_depth: self.inner.config.max_depth as i32,   // config.preview
            seed: self.inner.seed, world: 1, ..Default::default()
        };
        let cancel = CancelFlag::watch(rd.cancel_event);
        let mut rgb = vec![0f64; w * h * 3];
        check(unsafe { sys::rc_render_preview(ctx, &params, self.scale_width as i32, self.scale_height as i32,
                                              rgb.as_mut_ptr(), cancel.as_ptr()) })?;
        if cancel.is_set() { return Ok(()); }
        let px: Vec<Vec3> = rgb.chunks_exact(3).map(|c| Vec3::new(c[0], c[1], c[2])).collect();
        writer.write(ImageBufferEvent::BufferUpdate { rgb: px, r: 0, c: 0, width: w, height: h })
    }
}

impl Drop for CudaRenderer {
    fn drop(&mut self) { unsafe { sys::rc_destroy(*self.ctx.lock().unwrap()); } }
}

/// Bridges `Option<&SignalEvent>` (src/renderer.rs:25-30) to the `const int*` the C ABI polls.
pub struct CancelFlag { flag: std::sync::Arc<std::sync::atomic::AtomicI32> }
impl CancelFlag {
    pub fn watch(_event: Option<&SignalEvent>) -> Self { /* spawn a scoped poller that stores 1 when the event fires */
        Self { flag: Default::default() } }
    pub fn as_ptr(&self) -> *const i32 { self.flag.as_ptr() as *const i32 }
    pub fn is_set(&self) -> bool { self.flag.load(std::sync::atomic::Ordering::Relaxed) != 0 }
}

//! src/renderer/cuda.rs (new file) — `impl Renderer for CudaRenderer` / `CudaRendererScaled`: the drop-in behind
//! `Renderer::render` (src/renderer.rs:101-107) that hands the frame to libracer_cuda.so.
//!
//! * one `rc_ctx` per process, shared by the full and the preview renderer (both are built by the factory in
//!   src/renderer.rs:109-116, and they never render at the same time: src/scene_controller/interactive.rs:196-267);
//! * the scene is flattened (src/flatten.rs) and uploaded when `Hittable::generation()` has changed — the BVH bumps
//!   it whenever it was rebuilt (src/bvh_node.rs:176-205, the `bvh.changed()` of src/main.rs:178-183);
//! * the camera travels with every render call (14 f64 fields, src/camera.rs:57-72);
//! * `RenderData.cancel_event` (src/renderer.rs:25-30) is bridged to the `const int32_t*` the C ABI polls by a
//!   scoped thread that waits on the event; a cancelled render returns Ok(()) without writing (cpu.rs:55-62);
//! * the image comes back as `sqrt(mean)` per channel in f64, row 0 on top, NOT tone-mapped — what
//!   `ImageBufferEvent::BufferUpdate` carries (src/image_buffer.rs:63-71) — and is sent in a few large messages
//!   (the bus holds 1024 messages and a full bus is fatal: src/data_bus.rs);
//! * full renders pass `specialize: 2` (the scene compiled into the kernel when NVRTC is available, else the
//!   precompiled kernel); previews pass 0 (no run-time compilation while the user drags the camera).
//! There is no CPU fallback: without a usable device every render returns `TracerError::CudaBackend`.
use std::sync::atomic::{AtomicBool, AtomicI32, Ordering};
use std::sync::{Mutex, OnceLock};
use std::time::Duration;

use racer_cuda_sys as sys;
use synchronoise::SignalEvent;

use crate::{
    config::RenderConfig,
    data_bus::DataWriter,
    error::TracerError,
    flatten::FlatScene,
    image::Image,
    image_buffer::ImageBufferEvent,
    renderer::{RenderData, Renderer},
    vec3::Vec3,
};

struct CudaState {
    ctx: *mut sys::rc_ctx,
    uploaded: Option<u64>, // generation of the scene the device holds
    seed: u64,
}
// the handle is used by one thread at a time (behind the mutex); the library calls cudaSetDevice itself
unsafe impl Send for CudaState {}

fn state() -> &'static Mutex<CudaState> {
    static STATE: OnceLock<Mutex<CudaState>> = OnceLock::new();
    STATE.get_or_init(|| {
        Mutex::new(CudaState {
            ctx: std::ptr::null_mut(),
            uploaded: None,
            // the reference draws from an OS-seeded generator (src/util.rs:9-23); RACER_CUDA_SEED makes a run repeatable
            seed: std::env::var("RACER_CUDA_SEED").ok().and_then(|s| s.parse().ok()).unwrap_or_else(rand::random),
        })
    })
}

fn check(status: i32) -> Result<(), TracerError> {
    if status == sys::RC_OK {
        return Ok(());
    }
    if status == sys::RC_ERR_CANCELLED {
        return Err(TracerError::CancelEvent); // cancelled before it started, src/renderer/cpu.rs:79-83
    }
    let msg = unsafe { std::ffi::CStr::from_ptr(sys::rc_last_error()) }.to_string_lossy().into_owned();
    Err(TracerError::CudaBackend(status, msg))
}

/// Devices of this process: RACER_CUDA_DEVICES="0,1,2,3" (tile split over them with peer stores), default device 0.
fn devices() -> Vec<i32> {
    std::env::var("RACER_CUDA_DEVICES")
        .ok()
        .map(|s| s.split(',').filter_map(|d| d.trim().parse().ok()).collect::<Vec<i32>>())
        .filter(|v| !v.is_empty())
        .unwrap_or_else(|| vec![0])
}

fn with_ctx<R>(f: impl FnOnce(&mut CudaState) -> Result<R, TracerError>) -> Result<R, TracerError> {
    let mut guard = state().lock().map_err(|_| TracerError::FailedToAcquireLock("cuda context".to_string()))?;
    if guard.ctx.is_null() {
        let devs = devices();
        let mut ctx = std::ptr::null_mut();
        check(unsafe { sys::rc_create(devs.as_ptr(), devs.len() as i32, &mut ctx) })?;
        guard.ctx = ctx;
    }
    f(&mut guard)
}

/// Flatten + upload when the scene the device holds is not the one being rendered.
fn sync_scene(st: &mut CudaState, rd: &RenderData) -> Result<(), TracerError> {
    let generation = rd.scene.generation();
    if generation != 0 && st.uploaded == Some(generation) {
        return Ok(());
    }
    let mut flat = FlatScene::default();
    rd.scene.flatten(&mut flat)?;
    rd.background.flatten(&mut flat);
    if flat.n_prims() == 0 {
        return Err(TracerError::CudaBackend(-1, "the scene did not describe itself (Hittable::flatten)".to_string()));
    }
    let images: Vec<sys::rc_image> = flat
        .images
        .iter()
        .map(|(w, h, px)| sys::rc_image { width: *w, height: *h, rgba: px.as_ptr() })
        .collect();
    let scene = flat.as_rc_scene(&images);
    check(unsafe { sys::rc_upload_scene(st.ctx, &scene) })?;
    st.uploaded = Some(generation);
    Ok(())
}

/// Runs `f` with the `const int32_t*` the library polls; a scoped thread raises it when `event` fires.
/// Returns f's result and whether the flag was raised.
fn with_cancel_flag<R>(event: Option<&SignalEvent>, f: impl FnOnce(*const i32) -> R) -> (R, bool) {
    let flag = AtomicI32::new(0);
    let done = AtomicBool::new(false);
    let result = std::thread::scope(|s| {
        if let Some(ev) = event {
            let (flag, done) = (&flag, &done);
            s.spawn(move || {
                while !done.load(Ordering::Acquire) {
                    if ev.wait_timeout(Duration::from_millis(1)) {
                        flag.store(1, Ordering::Release);
                        break;
                    }
                }
            });
        }
        let r = f(if event.is_some() { flag.as_ptr() as *const i32 } else { std::ptr::null() });
        done.store(true, Ordering::Release);
        r
    });
    (result, flag.load(Ordering::Acquire) != 0)
}

fn params(st: &CudaState, image: &Image, config: &RenderConfig, specialize: i32) -> sys::rc_params {
    sys::rc_params {
        width: image.width as i32,
        height: image.height as i32,
        samples: config.samples as i32,     // src/renderer/cpu.rs:38
        max_depth: config.max_depth as i32, // src/renderer/cpu.rs:46
        seed: st.seed,
        variant: sys::RC_VARIANT_MEGAKERNEL,
        sampler: sys::RC_SAMPLER_DIRECT,
        split: sys::RC_SPLIT_TILES,
        world: 1,
        specialize,
        ..Default::default()
    }
}

/// Few, large `BufferUpdate` messages: bands of rows, like the reference's tiles but 16 of them instead of 100.
fn send(rgb: &[f64], width: usize, height: usize, writer: &DataWriter<ImageBufferEvent>) -> Result<(), TracerError> {
    let band = (height + 15) / 16;
    let mut r = 0;
    while r < height {
        let rows = band.min(height - r);
        let px: Vec<Vec3> = rgb[r * width * 3..(r + rows) * width * 3]
            .chunks_exact(3)
            .map(|c| Vec3::new(c[0], c[1], c[2]))
            .collect();
        writer.write(ImageBufferEvent::BufferUpdate { rgb: px, r, c: 0, width, height: rows })?;
        r += rows;
    }
    Ok(())
}

/// `renderer: Cuda` — the drop-in for `CpuRenderer` (src/renderer/cpu.rs).
pub struct CudaRenderer {
    config: RenderConfig,
}

impl CudaRenderer {
    pub fn new(config: RenderConfig) -> Self {
        Self { config }
    }
}

impl Renderer for CudaRenderer {
    fn render(&self, rd: RenderData, writer: &DataWriter<ImageBufferEvent>) -> Result<(), TracerError> {
        let (w, h) = (rd.image.width, rd.image.height);
        let mut rgb = vec![0f64; w * h * 3];
        let cancelled = with_ctx(|st| {
            sync_scene(st, &rd)?;
            check(unsafe { sys::rc_set_camera(st.ctx, &rd.camera_data.to_rc_camera()) })?;
            // 2: the scene compiled into the kernel (NVRTC) when possible, else the precompiled kernel
            let p = params(st, rd.image, &self.config, 2);
            st.seed = st.seed.wrapping_add(1); // a new stream per frame, as an unseeded generator would give
            let (status, cancelled) = with_cancel_flag(rd.cancel_event, |flag| unsafe {
                sys::rc_render(st.ctx, &p, rgb.as_mut_ptr(), flag)
            });
            check(status).map(|_| cancelled)
        })?;
        if cancelled {
            return Ok(()); // a cancelled render writes nothing, src/renderer/cpu.rs:55-62
        }
        send(&rgb, w, h, writer)
    }
}

fn get_highest_divdable(value: usize, mut div: usize) -> usize {
    // src/renderer/cpu_scaled.rs:17-23
    while (value % div) != 0 {
        div -= 1;
    }
    div
}

/// `preview_renderer: CudaPreview` — the drop-in for `CpuRendererScaled` (src/renderer/cpu_scaled.rs).
pub struct CudaRendererScaled {
    config: RenderConfig,
    scale_width: usize,
    scale_height: usize,
}

impl CudaRendererScaled {
    pub fn new(config: RenderConfig, image: &Image) -> Self {
        // src/renderer/cpu_scaled.rs:31-34
        let scale_width = get_highest_divdable(image.width / config.num_threads_width, config.scale);
        let scale_height = get_highest_divdable(image.height / config.num_threads_height, config.scale);
        Self { config, scale_width, scale_height }
    }
}

impl Renderer for CudaRendererScaled {
    fn render(&self, rd: RenderData, writer: &DataWriter<ImageBufferEvent>) -> Result<(), TracerError> {
        let (w, h) = (rd.image.width, rd.image.height);
        let mut rgb = vec![0f64; w * h * 3];
        let cancelled = with_ctx(|st| {
            sync_scene(st, &rd)?;
            check(unsafe { sys::rc_set_camera(st.ctx, &rd.camera_data.to_rc_camera()) })?;
            // config.preview: samples, max_depth, scale; 0: no run-time compilation on the interactive path
            let p = params(st, rd.image, &self.config, 0);
            st.seed = st.seed.wrapping_add(1);
            let (status, cancelled) = with_cancel_flag(rd.cancel_event, |flag| unsafe {
                sys::rc_render_preview(st.ctx, &p, self.scale_width as i32, self.scale_height as i32, rgb.as_mut_ptr(), flag)
            });
            check(status).map(|_| cancelled)
        })?;
        if cancelled {
            return Ok(());
        }
        send(&rgb, w, h, writer)
    }
}

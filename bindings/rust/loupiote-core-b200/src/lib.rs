//! `loupiote-core`-shaped wrapper over libloupiote_b200: same type and method names as the
//! reference crate (crates/lib/src/{lib,device,scene,renderer,errors}.rs and loaders/), so the
//! standalone application can switch its `use loupiote_core::…` lines to this crate.  The
//! wgpu arguments of the reference signatures (device, queue, encoder) have no counterpart:
//! the library owns its CUDA context and stream.  SOURCE ONLY -- no Rust toolchain exists in
//! the image this was written in; the tested host is the Python mirror (loupiote_b200/api.py)
//! and the C++ CLI (tools/cli/lp_render.cpp), which make exactly these calls.
use loupiote_b200_sys as ffi;
use std::ffi::{CStr, CString};
use std::os::raw::c_int;
use std::path::Path;
use std::ptr::{null, null_mut};

/// crates/lib/src/errors.rs:2-6 (+ the library's own failure classes as `Backend`).
#[derive(Debug)]
pub enum Error {
    FileNotFound(String),
    TextureToBufferReadFail,
    AccelBuild(String),
    Backend(c_int, String),
}

fn last_error() -> String {
    unsafe { CStr::from_ptr(ffi::lp_last_error()).to_string_lossy().into_owned() }
}

fn check(status: c_int) -> Result<(), Error> {
    match status {
        ffi::LP_OK => Ok(()),
        ffi::LP_ERR_FILE_NOT_FOUND => Err(Error::FileNotFound(last_error())),
        ffi::LP_ERR_READBACK => Err(Error::TextureToBufferReadFail),
        ffi::LP_ERR_ACCEL_BUILD => Err(Error::AccelBuild(last_error())),
        other => Err(Error::Backend(other, last_error())),
    }
}

/// device.rs:71-141.  `Device::new(cuda_ordinal)` replaces `Device::new(wgpu::Device)`.
pub struct Device {
    raw: *mut ffi::lp_device,
}
impl Device {
    pub fn new(cuda_ordinal: i32) -> Result<Self, Error> {
        let mut raw = null_mut();
        check(unsafe { ffi::lp_device_create(cuda_ordinal, &mut raw) })?;
        Ok(Device { raw })
    }
    pub fn synchronize(&self) -> Result<(), Error> {
        check(unsafe { ffi::lp_device_synchronize(self.raw) })
    }
}
impl Drop for Device {
    fn drop(&mut self) {
        unsafe { ffi::lp_device_destroy(self.raw) };
    }
}

pub use ffi::lp_light as Light;
pub use ffi::lp_material as Material;

/// albedo_rtx::MeshDescriptor (gltf.rs:91-95): strided views, `pas::Slice` = pointer + stride.
pub struct MeshDescriptor<'a> {
    pub positions: &'a [[f32; 4]],
    pub normals: Option<&'a [[f32; 3]]>,
    pub texcoords0: Option<&'a [[f32; 2]]>,
}
pub struct IndexedMeshDescriptor<'a> {
    pub mesh: MeshDescriptor<'a>,
    pub indices: &'a [u32],
}

/// scene.rs:5-28.
pub struct ImageData {
    pub data: Vec<u8>,
    pub width: u32,
    pub height: u32,
}

/// scene.rs:30-54 -- `Scene::default()` seeds the dummy index-0 entries.
pub struct Scene {
    raw: *mut ffi::lp_scene,
}
impl Default for Scene {
    fn default() -> Self {
        let mut raw = null_mut();
        check(unsafe { ffi::lp_scene_create(&mut raw) }).expect("lp_scene_create");
        Scene { raw }
    }
}
impl Drop for Scene {
    fn drop(&mut self) {
        unsafe { ffi::lp_scene_destroy(self.raw) };
    }
}
impl Scene {
    /// `scene.blas` of the reference is reached through these methods (BLASArray,
    /// gltf.rs:97-105,141-145).
    pub fn add_bvh(&mut self, mesh: MeshDescriptor) -> Result<u32, Error> {
        self.add(mesh, None)
    }
    pub fn add_bvh_indexed(&mut self, desc: IndexedMeshDescriptor) -> Result<u32, Error> {
        self.add(desc.mesh, Some(desc.indices))
    }
    fn add(&mut self, m: MeshDescriptor, indices: Option<&[u32]>) -> Result<u32, Error> {
        let mut out = 0u32;
        let n = m.normals.map_or(null(), |v| v.as_ptr() as *const _);
        let t = m.texcoords0.map_or(null(), |v| v.as_ptr() as *const _);
        let p = m.positions.as_ptr() as *const _;
        check(unsafe {
            match indices {
                None => ffi::lp_scene_add_bvh(self.raw, p, 16, n, 12, t, 8, m.positions.len(), &mut out),
                Some(i) => ffi::lp_scene_add_bvh_indexed(
                    self.raw, p, 16, n, 12, t, 8, m.positions.len(), i.as_ptr(), i.len(), &mut out),
            }
        })?;
        Ok(out)
    }
    pub fn add_instance(&mut self, blas: u32, model_to_world: glam::Mat4, material: u32) -> Result<(), Error> {
        check(unsafe {
            ffi::lp_scene_add_instance(self.raw, blas, model_to_world.to_cols_array().as_ptr(), material)
        })
    }
    /// `scene.blas.instances[i].set_transform(m)` (standalone/src/lib.rs:118-121).
    pub fn set_instance_transform(&mut self, instance: u32, m: glam::Mat4) -> Result<(), Error> {
        check(unsafe { ffi::lp_scene_set_instance_transform(self.raw, instance, m.to_cols_array().as_ptr()) })
    }
    pub fn push_material(&mut self, m: &Material) -> Result<u32, Error> {
        let mut out = 0u32;
        check(unsafe { ffi::lp_scene_push_material(self.raw, m, &mut out) })?;
        Ok(out)
    }
    pub fn push_light(&mut self, l: &Light) -> Result<u32, Error> {
        let mut out = 0u32;
        check(unsafe { ffi::lp_scene_push_light(self.raw, l, &mut out) })?;
        Ok(out)
    }
    pub fn push_image(&mut self, img: &ImageData) -> Result<u32, Error> {
        let mut out = 0u32;
        check(unsafe { ffi::lp_scene_push_image(self.raw, img.data.as_ptr(), img.width, img.height, &mut out) })?;
        Ok(out)
    }
    /// `scene.materials[i] = m` / `scene.lights[i] = l` (pub fields, scene.rs:30-35): edits of
    /// existing entries; `SceneGPU::update_instances` carries them to the device.
    pub fn set_material(&mut self, index: u32, m: &Material) -> Result<(), Error> {
        check(unsafe { ffi::lp_scene_set_material(self.raw, index, m) })
    }
    pub fn set_light(&mut self, index: u32, l: &Light) -> Result<(), Error> {
        check(unsafe { ffi::lp_scene_set_light(self.raw, index, l) })
    }
    /// Extension: a deforming mesh -- new positions (and normals) for the vertices of an
    /// existing BLAS; the tree is refitted, not rebuilt.  `SceneGPU::refit` follows.
    pub fn update_bvh_vertices(&mut self, blas: u32, mesh: MeshDescriptor) -> Result<(), Error> {
        let n = mesh.normals.map_or(null(), |v| v.as_ptr() as *const _);
        check(unsafe {
            ffi::lp_scene_update_bvh_vertices(
                self.raw, blas, mesh.positions.as_ptr() as *const _, 16, n, 12, mesh.positions.len())
        })
    }
}

/// crates/lib/src/loaders (gltf.rs:46-161, binary.rs:6-70).
pub mod loaders {
    use super::*;
    pub fn load_gltf(data: &[u8], scene: &mut Scene) -> Result<(), Error> {
        check(unsafe { ffi::lp_load_gltf(data.as_ptr(), data.len(), scene.raw) })
    }
    pub fn load_gltf_path<P: AsRef<Path>>(path: P, scene: &mut Scene) -> Result<(), Error> {
        let c = CString::new(path.as_ref().to_string_lossy().as_bytes()).unwrap();
        check(unsafe { ffi::lp_load_gltf_path(c.as_ptr(), scene.raw) })
    }
    pub fn load_binary_from_path<P: AsRef<Path>>(path: P, scene: &mut Scene) -> Result<(), Error> {
        let c = CString::new(path.as_ref().to_string_lossy().as_bytes()).unwrap();
        check(unsafe { ffi::lp_load_binary_from_path(c.as_ptr(), scene.raw) })
    }
}

/// scene.rs:56-64,151-187: TLAS build, re-layout, texture atlas, upload.
pub struct SceneGPU {
    raw: *mut ffi::lp_scene_gpu,
}
impl SceneGPU {
    pub fn new_from_scene(scene: &Scene, device: &Device) -> Result<Self, Error> {
        let mut raw = null_mut();
        check(unsafe { ffi::lp_scene_gpu_new_from_scene(scene.raw, device.raw, &mut raw) })?;
        Ok(SceneGPU { raw })
    }
    /// Extension: the same upload with every BLAS and the TLAS built on the device (LBVH),
    /// for scenes whose host SAH build would dominate the load time.
    pub fn new_from_scene_device_built(scene: &Scene, device: &Device) -> Result<Self, Error> {
        let mut raw = null_mut();
        check(unsafe { ffi::lp_scene_gpu_new_from_scene_lbvh(scene.raw, device.raw, &mut raw) })?;
        Ok(SceneGPU { raw })
    }
    /// After `Scene::set_instance_transform` / `set_material` / `set_light`: TLAS region,
    /// instance records and small tables only.
    pub fn update_instances(&mut self, scene: &Scene) -> Result<(), Error> {
        check(unsafe { ffi::lp_scene_gpu_update_instances(self.raw, scene.raw) })
    }
    /// After `Scene::update_bvh_vertices`: in place, renderers keep their binding.
    pub fn refit(&mut self, scene: &Scene) -> Result<(), Error> {
        check(unsafe { ffi::lp_scene_gpu_refit(self.raw, scene.raw) })
    }
}
impl Drop for SceneGPU {
    fn drop(&mut self) {
        unsafe { ffi::lp_scene_gpu_destroy(self.raw) };
    }
}

/// scene.rs:66-121: RGBE8 equirect probe (importance sampled by the library).
pub struct ProbeGPU {
    raw: *mut ffi::lp_probe,
}
impl ProbeGPU {
    pub fn new(device: &Device, data: &[u8], width: u32, height: u32) -> Result<Self, Error> {
        assert_eq!(data.len(), (width * height * 4) as usize);
        let mut raw = null_mut();
        check(unsafe { ffi::lp_probe_new(device.raw, data.as_ptr(), width, height, &mut raw) })?;
        Ok(ProbeGPU { raw })
    }
}
impl Drop for ProbeGPU {
    fn drop(&mut self) {
        unsafe { ffi::lp_probe_destroy(self.raw) };
    }
}

/// renderer.rs:160-167 (spelling of the first variant kept).
#[derive(Clone, Copy, PartialEq, Eq)]
pub enum BlitMode {
    Pahtrace = 0,
    DenoisedPathrace = 1,
    Temporal = 2,
    GBuffer = 3,
    MotionVector = 4,
}

/// renderer.rs:169-811.  The pub fields `accumulate` / `downsample_factor` are pushed to the
/// library at the call that consumes them, as the reference reads them (renderer.rs:203-204).
pub struct Renderer {
    raw: *mut ffi::lp_renderer,
    size: (u32, u32),
    pub accumulate: bool,
    pub downsample_factor: f32,
}
impl Renderer {
    pub fn max_ssbo_element_in_bytes() -> u32 {
        unsafe { ffi::lp_renderer_max_ssbo_element_in_bytes() }
    }
    pub fn new(device: &Device, original_size: (u32, u32)) -> Result<Self, Error> {
        let mut raw = null_mut();
        check(unsafe { ffi::lp_renderer_new(device.raw, original_size.0, original_size.1, &mut raw) })?;
        let mut r = Renderer { raw, size: (0, 0), accumulate: false, downsample_factor: 0.5 };
        r.refresh_size();
        Ok(r)
    }
    fn refresh_size(&mut self) {
        unsafe { ffi::lp_renderer_get_size(self.raw, &mut self.size.0, &mut self.size.1) };
    }
    pub fn resize(&mut self, scene: &SceneGPU, probe: Option<&ProbeGPU>, size: (u32, u32)) -> Result<(), Error> {
        check(unsafe { ffi::lp_renderer_set_downsample_factor(self.raw, self.downsample_factor) })?;
        check(unsafe {
            ffi::lp_renderer_resize(self.raw, scene.raw, probe.map_or(null_mut(), |p| p.raw), size.0, size.1)
        })?;
        self.refresh_size();
        Ok(())
    }
    /// The renderer borrows `scene` / `probe` until the next set_resources / resize
    /// (renderer.rs:356,704-724 has the same lifetime through its bind groups).
    pub fn set_resources(&mut self, scene: &SceneGPU, probe: Option<&ProbeGPU>) -> Result<(), Error> {
        check(unsafe { ffi::lp_renderer_set_resources(self.raw, scene.raw, probe.map_or(null_mut(), |p| p.raw)) })
    }
    pub fn raytrace(&mut self, view_transform: &glam::Mat4) -> Result<(), Error> {
        check(unsafe { ffi::lp_renderer_set_accumulate(self.raw, self.accumulate as c_int) })?;
        check(unsafe { ffi::lp_renderer_raytrace(self.raw, view_transform.to_cols_array().as_ptr()) })
    }
    pub fn reset_accumulation(&mut self) -> Result<(), Error> {
        self.accumulate = false;
        check(unsafe { ffi::lp_renderer_reset_accumulation(self.raw) })
    }
    pub fn upload_noise_texture(&mut self, data: &[u8], width: u32, height: u32, bytes_per_row: u32) -> Result<(), Error> {
        check(unsafe { ffi::lp_renderer_upload_noise_texture(self.raw, data.as_ptr(), width, height, bytes_per_row) })
    }
    pub fn use_noise_texture(&mut self, flag: bool) -> Result<(), Error> {
        check(unsafe { ffi::lp_renderer_use_noise_texture(self.raw, flag as c_int) })
    }
    pub fn set_blit_mode(&mut self, mode: BlitMode) -> Result<(), Error> {
        check(unsafe { ffi::lp_renderer_set_blit_mode(self.raw, mode as c_int) })
    }
    pub fn get_size(&self) -> &(u32, u32) {
        &self.size
    }
    /// renderer.rs:727-811: sRGB8 bytes of the main target, `w * h * 4`.
    pub fn read_pixels(&self) -> Result<Vec<u8>, Error> {
        let mut out = vec![0u8; (self.size.0 * self.size.1 * 4) as usize];
        check(unsafe { ffi::lp_renderer_read_pixels(self.raw, out.as_mut_ptr(), out.len()) })?;
        Ok(out)
    }
    /// gpu::Queries labels()/values() (renderer.rs:444-517, performance_info.rs:19-20).
    pub fn queries(&mut self) -> Result<Vec<(String, f64)>, Error> {
        let (mut labels, mut ms, mut n) = (null(), null(), 0usize);
        check(unsafe { ffi::lp_renderer_queries(self.raw, &mut labels, &mut ms, &mut n) })?;
        Ok((0..n)
            .map(|i| unsafe {
                (CStr::from_ptr(*labels.add(i)).to_string_lossy().into_owned(), *ms.add(i))
            })
            .collect())
    }
    /// Knobs that are constants in the reference (bounces, spp per frame, a-trous count).
    pub fn set_config(&mut self, f: impl FnOnce(&mut ffi::lp_render_config)) -> Result<(), Error> {
        let mut cfg = unsafe { std::mem::zeroed::<ffi::lp_render_config>() };
        check(unsafe { ffi::lp_renderer_get_config(self.raw, &mut cfg) })?;
        f(&mut cfg);
        check(unsafe { ffi::lp_renderer_set_config(self.raw, &cfg) })
    }
}
impl Drop for Renderer {
    fn drop(&mut self) {
        unsafe { ffi::lp_renderer_destroy(self.raw) };
    }
}

/// Extension (the reference renders on one device, standalone/src/lib.rs:220-231): the Renderer
/// on several GPUs of one box.  The scene is replicated, the samples of a frame are split --
/// rank g of W traces sample indices g, g+W, ... of the sequence one GPU would trace -- and the
/// FP32 accumulators are summed to rank 0 (NCCL inside the library), where the tone map follows.
pub struct MultiRenderer {
    raw: *mut ffi::lp_multi,
    size: (u32, u32),
}
impl MultiRenderer {
    /// ONE process drives the GPUs `ordinals`.
    pub fn new(ordinals: &[i32]) -> Result<Self, Error> {
        let mut raw = null_mut();
        check(unsafe { ffi::lp_multi_create(ordinals.as_ptr(), ordinals.len() as c_int, &mut raw) })?;
        Ok(MultiRenderer { raw, size: (0, 0) })
    }
    /// One process per GPU: rank 0 calls `unique_id` and carries the bytes to the others.
    pub fn unique_id() -> Result<[u8; 128], Error> {
        let mut id = [0u8; 128];
        check(unsafe { ffi::lp_multi_unique_id(id.as_mut_ptr()) })?;
        Ok(id)
    }
    pub fn new_rank(ordinal: i32, id: &[u8; 128], world: i32, rank: i32) -> Result<Self, Error> {
        let mut raw = null_mut();
        check(unsafe { ffi::lp_multi_create_rank(ordinal, id.as_ptr(), world, rank, &mut raw) })?;
        Ok(MultiRenderer { raw, size: (0, 0) })
    }
    /// `SceneGPU::new_from_scene` + `Renderer::set_resources` on every GPU.
    pub fn set_scene(&mut self, scene: &Scene, device_build: bool) -> Result<(), Error> {
        check(unsafe { ffi::lp_multi_set_scene(self.raw, scene.raw, device_build as c_int) })
    }
    pub fn resize(&mut self, size: (u32, u32), downsample_factor: f32) -> Result<(), Error> {
        check(unsafe { ffi::lp_multi_resize(self.raw, size.0, size.1, downsample_factor) })?;
        self.size = ((size.0 as f32 * downsample_factor) as u32, (size.1 as f32 * downsample_factor) as u32);
        Ok(())
    }
    /// `cfg.spp_per_call` = total samples per pixel of one `raytrace` over all GPUs.
    pub fn set_config(&mut self, cfg: &ffi::lp_render_config) -> Result<(), Error> {
        check(unsafe { ffi::lp_multi_set_config(self.raw, cfg) })
    }
    pub fn set_accumulate(&mut self, flag: bool) -> Result<(), Error> {
        check(unsafe { ffi::lp_multi_set_accumulate(self.raw, flag as c_int) })
    }
    /// `Renderer::raytrace` on every GPU (asynchronous).
    pub fn raytrace(&mut self, view_transform: &glam::Mat4) -> Result<(), Error> {
        check(unsafe { ffi::lp_multi_render(self.raw, view_transform.to_cols_array().as_ptr()) })
    }
    /// The exchange step: accumulators summed to rank 0, tone map behind it (asynchronous).
    pub fn reduce(&mut self) -> Result<(), Error> {
        check(unsafe { ffi::lp_multi_reduce(self.raw) })
    }
    /// `Renderer::read_pixels` of the reduced frame (rank 0).
    pub fn read_pixels(&self) -> Result<Vec<u8>, Error> {
        let mut out = vec![0u8; (self.size.0 * self.size.1 * 4) as usize];
        check(unsafe { ffi::lp_multi_read_pixels(self.raw, out.as_mut_ptr(), out.len()) })?;
        Ok(out)
    }
}
impl Drop for MultiRenderer {
    fn drop(&mut self) {
        unsafe { ffi::lp_multi_destroy(self.raw) };
    }
}

// cuda_engine.rs -- the reference-side binding of libvello_b200.so (drop next to vello/src/wgpu_engine.rs).
//
// UNCOMPILED SOURCE: this image has no Rust toolchain (`cargo --version` fails; wgpu / kurbo / peniko are not vendored), so the
// file documents, in the reference's own language, exactly what a vello maintainer adds. Every `extern "C"` item below is
// declared in include/vello_b200.h; the struct layouts are checked against that header by tests/test_abi.py on the C side.
//
// What it replaces:  Render::render_encoding_coarse + record_fine   vello/src/render.rs:135-629
//                    WgpuEngine::run_recording                      vello/src/wgpu_engine.rs:380-780
// What it keeps:     Scene, Renderer, RenderParams, AaConfig, Resolver (or its device twin, see `render_streams`).
#![allow(non_camel_case_types, dead_code)]
use std::os::raw::{c_char, c_int, c_void};

use vello_encoding::{Encoding, Layout, Resolver};

#[repr(C)]
pub struct VbOptions { pub device: i32, pub timing: u32, pub max_retries: u32, pub reserved: u32 }

/// == RenderParams (vello/src/lib.rs:357-369) + the stripe window extension
#[repr(C)]
#[derive(Default, Clone, Copy)]
pub struct VbParams {
    pub base_color: u32, pub width: u32, pub height: u32, pub aa: u32,
    pub bin_row0: u32, pub bin_row1: u32, pub tile_row0: u32, pub tile_row1: u32,
}

#[repr(C)]
#[derive(Default)]
pub struct VbFrameStats {
    pub failed: u32, pub binning: u32, pub ptcl: u32, pub tile: u32, pub seg_counts: u32, pub segments: u32, pub blend: u32, pub lines: u32,
    pub retries: u32, pub kernel_launches: u32, pub stage_ms: [f32; 11], pub total_ms: f32, pub arena_bytes: u64,
}

#[repr(C)] pub struct VbRampStop { pub offset: f32, pub r: f32, pub g: f32, pub b: f32, pub a: f32 }
#[repr(C)] pub struct VbRampPatch { pub draw_data_offset: u32, pub extend: u32, pub premul_interp: u32, pub n_stops: u32, pub stops: *const VbRampStop }
#[repr(C)] pub struct VbImagePatch { pub draw_data_offset: u32, pub width: u32, pub height: u32, pub pixels: *const u8 }
/// == the six streams of vello_encoding::Encoding (encoding.rs:22-48) + its patches (resolve.rs:560-590)
#[repr(C)]
pub struct VbEncodingStreams {
    pub path_tags: *const u8, pub n_path_tags: u32,
    pub path_data: *const u32, pub n_path_data: u32,
    pub draw_tags: *const u32, pub n_draw_tags: u32,
    pub draw_data: *const u32, pub n_draw_data: u32,
    pub transforms: *const f32, pub n_transforms: u32,
    pub styles: *const u32, pub n_styles: u32,
    pub n_paths: u32, pub n_clips: u32, pub n_open_clips: u32,
    pub ramp_patches: *const VbRampPatch, pub n_ramp_patches: u32,
    pub image_patches: *const VbImagePatch, pub n_image_patches: u32,
}

pub enum VbRenderer {}
pub enum VbGroup {}

#[link(name = "vello_b200")]
extern "C" {
    pub fn vb_renderer_new(opt: *const VbOptions, out: *mut *mut VbRenderer) -> c_int;
    pub fn vb_renderer_free(r: *mut VbRenderer);
    pub fn vb_strerror(code: c_int) -> *const c_char;
    pub fn vb_last_error(r: *mut VbRenderer) -> *const c_char;
    // vello_encoding::Layout is #[repr(C)] 10 x u32 (resolve.rs:16-39) == vb_layout: passed by pointer.
    pub fn vb_render(r: *mut VbRenderer, scene: *const u8, scene_len: usize, layout: *const Layout,
                     ramps: *const u32, ramp_w: u32, ramp_h: u32, atlas: *const u8, atlas_w: u32, atlas_h: u32,
                     params: *const VbParams, out: *mut c_void, out_is_device: u32, stats: *mut VbFrameStats) -> c_int;
    // three frames in flight for exporters that read every frame back (examples/headless/src/main.rs:188-210)
    pub fn vb_render_begin(r: *mut VbRenderer, scene: *const u8, scene_len: usize, layout: *const Layout,
                           ramps: *const u32, ramp_w: u32, ramp_h: u32, atlas: *const u8, atlas_w: u32, atlas_h: u32,
                           params: *const VbParams, out_host: *mut c_void, stats: *mut VbFrameStats) -> c_int;
    pub fn vb_readback_wait(r: *mut VbRenderer) -> c_int;
    // Resolver::resolve on the device: hand over the raw streams instead of a packed buffer
    pub fn vb_scene_upload_streams(r: *mut VbRenderer, e: *const VbEncodingStreams, layout_out: *mut Layout) -> c_int;
    pub fn vb_render_uploaded(r: *mut VbRenderer, params: *const VbParams, out: *mut c_void, out_is_device: u32, stats: *mut VbFrameStats) -> c_int;
    pub fn vb_set_occlusion_cull(r: *mut VbRenderer, on: c_int) -> c_int;
    // one frame on several GPUs of one box
    pub fn vb_group_new(devices: *const i32, n: u32, opt: *const VbOptions, out: *mut *mut VbGroup) -> c_int;
    pub fn vb_group_free(g: *mut VbGroup);
    pub fn vb_group_render(g: *mut VbGroup, scene: *const u8, scene_len: usize, layout: *const Layout,
                           ramps: *const u32, ramp_w: u32, ramp_h: u32, atlas: *const u8, atlas_w: u32, atlas_h: u32,
                           params: *const VbParams, out: *mut c_void, out_is_device: u32, stats: *mut VbFrameStats) -> c_int;
}

#[derive(Debug)]
pub enum CudaError { Code(c_int) }

/// Sits where `WgpuEngine` sits: owned by `Renderer`, `Send` but not `Sync` (one host thread per renderer, lib.rs:351-352).
pub struct CudaEngine { raw: *mut VbRenderer, packed: Vec<u8> }
unsafe impl Send for CudaEngine {}

impl CudaEngine {
    /// `Renderer::new` (lib.rs:432-458). `RendererOptions::use_cpu` has no equivalent: there is no CPU fallback.
    pub fn new(device: i32) -> Result<Self, CudaError> {
        let opt = VbOptions { device, timing: 0, max_retries: 0, reserved: 0 };
        let mut raw = std::ptr::null_mut();
        let rc = unsafe { vb_renderer_new(&opt, &mut raw) };
        if rc != 0 { return Err(CudaError::Code(rc)); }
        Ok(Self { raw, packed: Vec::new() })
    }

    fn params(p: &crate::RenderParams) -> VbParams {
        VbParams {
            base_color: p.base_color.premultiply().to_rgba8().to_u32(), // config.rs:183
            width: p.width, height: p.height,
            aa: match p.antialiasing_method { crate::AaConfig::Area => 0, crate::AaConfig::Msaa8 => 1, crate::AaConfig::Msaa16 => 2 },
            ..Default::default()
        }
    }

    /// Body of `Renderer::render_to_texture` (lib.rs:474-515) with the host-side resolve kept as it is (render.rs:148).
    pub fn render(&mut self, resolver: &mut Resolver, encoding: &Encoding, params: &crate::RenderParams,
                  atlas_rgba8: &[u8], atlas_w: u32, atlas_h: u32, out_rgba8: &mut [u8]) -> Result<VbFrameStats, CudaError> {
        let (layout, ramps, _images) = resolver.resolve(encoding, &mut self.packed);
        let p = Self::params(params);
        let mut stats = VbFrameStats::default();
        let rc = unsafe {
            vb_render(self.raw, self.packed.as_ptr(), self.packed.len(), &layout, ramps.data.as_ptr(), ramps.width, ramps.height,
                      atlas_rgba8.as_ptr(), atlas_w, atlas_h, &p, out_rgba8.as_mut_ptr().cast(), 0, &mut stats)
        };
        // an arena overflow is grown-and-re-run inside the call; only a persisting one is an error (the reference leaves the
        // texture unwritten and says nothing, lib.rs:753-763)
        if rc != 0 { Err(CudaError::Code(rc)) } else { Ok(stats) }
    }

    /// Same frame without `Resolver::resolve` on the host: the streams go to the device as they are and are resolved there.
    pub fn render_streams(&mut self, e: &VbEncodingStreams, params: &crate::RenderParams, out_rgba8: &mut [u8]) -> Result<VbFrameStats, CudaError> {
        let mut layout = Layout::default();
        let rc = unsafe { vb_scene_upload_streams(self.raw, e, &mut layout) };
        if rc != 0 { return Err(CudaError::Code(rc)); }
        let p = Self::params(params);
        let mut stats = VbFrameStats::default();
        let rc = unsafe { vb_render_uploaded(self.raw, &p, out_rgba8.as_mut_ptr().cast(), 0, &mut stats) };
        if rc != 0 { Err(CudaError::Code(rc)) } else { Ok(stats) }
    }
}

impl Drop for CudaEngine {
    fn drop(&mut self) { unsafe { vb_renderer_free(self.raw) } }
}

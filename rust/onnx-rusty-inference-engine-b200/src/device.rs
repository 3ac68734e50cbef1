//! Safe handles over b200rt-sys: one `Context` per GPU, `DeviceTensor` = a value of the store (replaces the host
//! `Array2<f32>` / `Array4<f32>` of model_inference.rs:30-32), `Model` = the graph-level executor.
use std::collections::HashMap;
use std::ffi::{CStr, CString};
use std::ptr;
use std::sync::{Arc, Mutex, OnceLock};

use b200rt_sys as sys;
use ndarray::{Array2, Array4};
use onnx_protobuf::TensorProto;

/// name -> (2-D slot, 4-D slot), exactly the reference's store with device handles as values.
pub type Store = Arc<Mutex<HashMap<String, (Option<DeviceTensor>, Option<DeviceTensor>)>>>;

#[derive(Debug)]
pub struct B200Error { pub code: i32, pub message: String }

pub fn last_error(code: i32) -> B200Error {
    let message = unsafe { CStr::from_ptr(sys::b200_last_error()) }.to_string_lossy().into_owned();
    B200Error { code, message }
}

pub fn check(rc: i32) -> Result<(), B200Error> { if rc == 0 { Ok(()) } else { Err(last_error(rc)) } }

struct CtxInner(*mut sys::b200_ctx);
unsafe impl Send for CtxInner {}
unsafe impl Sync for CtxInner {}
impl Drop for CtxInner { fn drop(&mut self) { unsafe { sys::b200_ctx_destroy(self.0); } } }

/// One per GPU; calls on one context are stream-ordered and serialised by the library's internal lock, like the
/// reference's mutex-guarded store.
#[derive(Clone)]
pub struct Context(Arc<CtxInner>);

impl Context {
    pub fn new(device: i32) -> Result<Context, B200Error> {
        let mut p = ptr::null_mut();
        check(unsafe { sys::b200_ctx_create(device, ptr::null_mut(), &mut p) })?;   // fails without an sm_100 GPU: no CPU fallback
        Ok(Context(Arc::new(CtxInner(p))))
    }
    pub fn raw(&self) -> *mut sys::b200_ctx { self.0 .0 }
    pub fn sync(&self) -> Result<(), B200Error> { check(unsafe { sys::b200_sync(self.raw()) }) }
    pub fn upload(&self, dims: &[i64], data: &[f32]) -> Result<DeviceTensor, B200Error> {
        let mut t = ptr::null_mut();
        check(unsafe { sys::b200_tensor_alloc(self.raw(), dims.as_ptr(), dims.len() as i32, &mut t) })?;
        let t = DeviceTensor::from_raw(self.clone(), t);
        check(unsafe { sys::b200_tensor_upload(t.raw(), data.as_ptr(), data.len()) })?;
        Ok(t)
    }
}

/// The process-wide context the reference-shaped functions use (their signatures carry none): device 0.
pub fn default_context() -> &'static Context {
    static CTX: OnceLock<Context> = OnceLock::new();
    CTX.get_or_init(|| Context::new(0).unwrap_or_else(|e| panic!("b200rt: {}", e.message)))
}

struct TensorInner { ctx: Context, t: *mut sys::b200_tensor }
unsafe impl Send for TensorInner {}
unsafe impl Sync for TensorInner {}
impl Drop for TensorInner { fn drop(&mut self) { unsafe { sys::b200_tensor_free(self.t); } let _ = &self.ctx; } }

/// An HBM-resident fp32 tensor.  `clone()` shares the handle (the reference deep-clones its ndarrays on every read).
#[derive(Clone)]
pub struct DeviceTensor(Arc<TensorInner>);

impl DeviceTensor {
    pub fn from_raw(ctx: Context, t: *mut sys::b200_tensor) -> DeviceTensor { DeviceTensor(Arc::new(TensorInner { ctx, t })) }
    pub fn raw(&self) -> *mut sys::b200_tensor { self.0.t }
    pub fn context(&self) -> &Context { &self.0.ctx }
    pub fn rank(&self) -> i32 { unsafe { sys::b200_tensor_rank(self.raw()) } }
    pub fn dims(&self) -> Vec<i64> {
        let mut d = [0i64; 4];
        unsafe { sys::b200_tensor_dims(self.raw(), d.as_mut_ptr()); }
        d[..self.rank() as usize].to_vec()
    }
    pub fn download(&self) -> Result<Vec<f32>, B200Error> {
        let n: i64 = self.dims().iter().product();
        let mut v = vec![0f32; n as usize];
        check(unsafe { sys::b200_tensor_download(self.raw(), v.as_mut_ptr(), v.len()) })?;
        Ok(v)
    }
    /// Compatibility shims: what the reference's store held.
    pub fn to_array2(&self) -> Array2<f32> {
        let d = self.dims();
        Array2::from_shape_vec((d[0] as usize, d[1] as usize), self.download().expect("download")).unwrap()
    }
    pub fn to_array4(&self) -> Array4<f32> {
        let d = self.dims();
        Array4::from_shape_vec((d[0] as usize, d[1] as usize, d[2] as usize, d[3] as usize), self.download().expect("download")).unwrap()
    }
}

/// Initializers are uploaded once per process and TensorProto (the reference re-decodes them on every use, utils.rs:113).
pub fn cached_initializer(t: &TensorProto, dims: &[i64], data: impl FnOnce() -> Vec<f32>) -> DeviceTensor {
    static CACHE: OnceLock<Mutex<HashMap<usize, DeviceTensor>>> = OnceLock::new();
    let key = t as *const TensorProto as usize;     // stable for the life of the Arc<ModelProto>
    let mut map = CACHE.get_or_init(|| Mutex::new(HashMap::new())).lock().unwrap();
    map.entry(key).or_insert_with(|| default_context().upload(dims, &data()).unwrap_or_else(|e| panic!("b200rt: {}", e.message))).clone()
}

struct ModelInner { _ctx: Context, m: *mut sys::b200_model }
unsafe impl Send for ModelInner {}
unsafe impl Sync for ModelInner {}
impl Drop for ModelInner { fn drop(&mut self) { unsafe { sys::b200_model_free(self.m); } } }

/// Graph-level executor (b200_model_*): weights uploaded once, Conv+Add+Relu / Concat / Dropout / Reshape fused or elided,
/// the node sequence replayed as one CUDA graph per batch size.
pub struct Model { inner: ModelInner, pub in_chw: [i64; 3], pub out_per_image: i64 }

impl Model {
    pub fn from_bytes(ctx: &Context, onnx: &[u8]) -> Result<Model, B200Error> {
        let mut m = ptr::null_mut();
        check(unsafe { sys::b200_model_load_onnx(ctx.raw(), onnx.as_ptr(), onnx.len(), &mut m) })?;
        let (mut chw, mut opi) = ([0i64; 3], 0i64);
        check(unsafe { sys::b200_model_io(m, chw.as_mut_ptr(), &mut opi) })?;
        Ok(Model { inner: ModelInner { _ctx: ctx.clone(), m }, in_chw: chw, out_per_image: opi })
    }
    pub fn raw(&self) -> *mut sys::b200_model { self.inner.m }
    pub fn set_option(&self, key: &str, value: i64) -> Result<(), B200Error> {
        let k = CString::new(key).unwrap();
        check(unsafe { sys::b200_model_set_option(self.raw(), k.as_ptr(), value) })
    }
    /// Host-to-host: `input` holds `batch` images (NCHW, dense); returns batch * out_per_image floats.
    pub fn run(&self, input: &[f32], batch: i64) -> Result<Vec<f32>, B200Error> {
        let mut out = vec![0f32; (batch * self.out_per_image) as usize];
        check(unsafe { sys::b200_model_run(self.raw(), input.as_ptr(), batch, out.as_mut_ptr()) })?;
        Ok(out)
    }
    /// Batch-sharded over several devices (one Model per Context / device), logits in image order.
    pub fn run_sharded(models: &[&Model], input: &[f32], batch: i64) -> Result<Vec<f32>, B200Error> {
        let raws: Vec<*mut sys::b200_model> = models.iter().map(|m| m.raw()).collect();
        let mut out = vec![0f32; (batch * models[0].out_per_image) as usize];
        check(unsafe { sys::b200_model_run_sharded(raws.as_ptr(), raws.len() as i32, input.as_ptr(), batch, out.as_mut_ptr()) })?;
        Ok(out)
    }
}

//! Safe handles over the C ABI: one context per (thread, GPU), RAII device matrices / trees / raw buffers.
use core::ffi::c_void;
use core::ptr;
use std::cell::RefCell;
use std::ffi::CStr;
use std::rc::Rc;

use b200zk_sys as sys;
use p3_matrix::dense::RowMajorMatrix;
use p3_matrix::Matrix;

use crate::{as_u32, as_u32_mut, F};

/// `B200ZK_ERR_*` code plus the library's message.
#[derive(Debug, Clone)]
pub struct Error {
    pub code: i32,
    pub message: String,
}
impl core::fmt::Display for Error {
    fn fmt(&self, f: &mut core::fmt::Formatter<'_>) -> core::fmt::Result {
        write!(f, "b200zk error {}: {}", self.code, self.message)
    }
}
impl std::error::Error for Error {}

/// One `b200zk_ctx`: device ordinal, stream, twiddle caches.  Not `Send`: the library wants one context per thread.
pub struct Ctx {
    pub(crate) raw: *mut sys::b200zk_ctx,
}
impl Ctx {
    /// Fails (there is no CPU fallback) when the device is missing or the library cannot initialise it.
    pub fn new(device: i32) -> Result<Rc<Self>, Error> {
        let mut raw = ptr::null_mut();
        let rc = unsafe { sys::b200zk_ctx_create(device, &mut raw) };
        if rc != 0 {
            return Err(Error { code: rc, message: format!("cannot create a context on CUDA device {device}") });
        }
        Ok(Rc::new(Ctx { raw }))
    }
    pub(crate) fn check(&self, rc: i32) -> Result<(), Error> {
        if rc == 0 {
            return Ok(());
        }
        let message = unsafe { CStr::from_ptr(sys::b200zk_last_error(self.raw)) }.to_string_lossy().into_owned();
        Err(Error { code: rc, message })
    }
    pub fn sync(&self) -> Result<(), Error> {
        self.check(unsafe { sys::b200zk_ctx_sync(self.raw) })
    }
    /// `RowMajorMatrix<F>` -> device (one H2D copy).
    pub fn upload(self: &Rc<Self>, m: &RowMajorMatrix<F>) -> Result<DeviceMatrix, Error> {
        let mut raw = ptr::null_mut();
        self.check(unsafe { sys::b200zk_mat_upload(self.raw, as_u32(&m.values), m.height() as u64, m.width() as u32, &mut raw) })?;
        Ok(DeviceMatrix { ctx: self.clone(), raw, owned: true })
    }
    /// Any `Matrix<F>` -> device.  Dense matrices are uploaded in place; views are materialised row-major first.
    pub fn upload_any<M: Matrix<F>>(self: &Rc<Self>, m: &M) -> Result<DeviceMatrix, Error> {
        let dense = RowMajorMatrix::new(m.rows().flatten().collect(), m.width());
        self.upload(&dense)
    }
}
impl Drop for Ctx {
    fn drop(&mut self) {
        unsafe { sys::b200zk_ctx_destroy(self.raw) }
    }
}

thread_local! {
    static CTX: RefCell<Option<Rc<Ctx>>> = const { RefCell::new(None) };
}
/// The calling thread's context on the GPU named by `B200ZK_DEVICE` (default 0), created on first use.  Rayon workers
/// each get their own: the library is re-entrant across contexts and every context has its own stream.
pub fn with_ctx<R>(f: impl FnOnce(&Rc<Ctx>) -> R) -> R {
    CTX.with(|slot| {
        let mut slot = slot.borrow_mut();
        if slot.is_none() {
            let dev = std::env::var("B200ZK_DEVICE").ok().and_then(|s| s.parse().ok()).unwrap_or(0);
            *slot = Some(Ctx::new(dev).expect("b200zk: no usable CUDA device (the hot path has no CPU fallback)"));
        }
        f(slot.as_ref().unwrap())
    })
}

/// Device-resident `RowMajorMatrix<F>`.
pub struct DeviceMatrix {
    pub(crate) ctx: Rc<Ctx>,
    pub(crate) raw: *mut sys::b200zk_mat,
    pub(crate) owned: bool,
}
impl DeviceMatrix {
    pub fn height(&self) -> usize {
        unsafe { sys::b200zk_mat_rows(self.raw) as usize }
    }
    pub fn width(&self) -> usize {
        unsafe { sys::b200zk_mat_width(self.raw) as usize }
    }
    /// D2H copy of the whole matrix.
    pub fn download(&self) -> Result<RowMajorMatrix<F>, Error> {
        let (h, w) = (self.height(), self.width());
        let mut values = F::zero_vec(h * w);
        self.ctx.check(unsafe { sys::b200zk_mat_download(self.ctx.raw, self.raw, as_u32_mut(&mut values)) })?;
        Ok(RowMajorMatrix::new(values, w))
    }
    /// D2H copy of `nrows` rows starting at `row0` (what an opening needs).
    pub fn download_rows(&self, row0: usize, nrows: usize) -> Result<Vec<F>, Error> {
        let mut values = F::zero_vec(nrows * self.width());
        self.ctx.check(unsafe { sys::b200zk_mat_download_rows(self.ctx.raw, self.raw, row0 as u64, nrows as u64, as_u32_mut(&mut values)) })?;
        Ok(values)
    }
    /// Give the handle away (to a tree that takes ownership).
    pub(crate) fn into_raw(mut self) -> *mut sys::b200zk_mat {
        self.owned = false;
        self.raw
    }
}
impl Drop for DeviceMatrix {
    fn drop(&mut self) {
        if self.owned && !self.raw.is_null() {
            unsafe { sys::b200zk_mat_free(self.ctx.raw, self.raw) }
        }
    }
}

/// Device-resident `MerkleTreeMmcs` prover data: the committed matrices and every digest layer.
pub struct Tree {
    pub(crate) ctx: Rc<Ctx>,
    pub(crate) raw: *mut sys::b200zk_tree,
}
impl Tree {
    pub fn depth(&self) -> usize {
        unsafe { sys::b200zk_tree_depth(self.raw) as usize }
    }
    pub fn num_matrices(&self) -> usize {
        unsafe { sys::b200zk_tree_num_mats(self.raw) as usize }
    }
    pub fn total_width(&self) -> usize {
        unsafe { sys::b200zk_tree_total_width(self.raw) as usize }
    }
    /// The root digest (`b200zk_tree_root`).  For a tree from `b200zk_lde_commit_host_async` this is the collection point:
    /// it waits for that commit only, not for commits issued after it.
    pub fn root(&self) -> Result<crate::Digest, Error> {
        let mut r = [F::default(); 8];
        self.ctx.check(unsafe { sys::b200zk_tree_root(self.ctx.raw, self.raw, r.as_mut_ptr() as *mut u32) })?;
        Ok(r)
    }
    /// Borrowed handle of committed matrix `i` (original order); lives as long as the tree.
    pub fn matrix(&self, i: usize) -> DeviceMatrix {
        let raw = unsafe { sys::b200zk_tree_mat(self.raw, i as u32) } as *mut sys::b200zk_mat;
        DeviceMatrix { ctx: self.ctx.clone(), raw, owned: false }
    }
}
impl Drop for Tree {
    fn drop(&mut self) {
        if !self.raw.is_null() {
            unsafe { sys::b200zk_tree_free(self.ctx.raw, self.raw) }
        }
    }
}

/// Raw device allocation from the library's stream-ordered pool (EF4 vectors of the open phase).
pub struct DeviceBuf {
    pub(crate) ctx: Rc<Ctx>,
    pub(crate) ptr: *mut c_void,
    pub bytes: usize,
}
impl DeviceBuf {
    pub fn new(ctx: &Rc<Ctx>, bytes: usize) -> Result<Self, Error> {
        let mut p = ptr::null_mut();
        ctx.check(unsafe { sys::b200zk_dev_alloc(ctx.raw, bytes as u64, &mut p) })?;
        Ok(DeviceBuf { ctx: ctx.clone(), ptr: p, bytes })
    }
    pub fn zeroed(ctx: &Rc<Ctx>, bytes: usize) -> Result<Self, Error> {
        let b = Self::new(ctx, bytes)?;
        ctx.check(unsafe { sys::b200zk_dev_zero(ctx.raw, b.ptr, bytes as u64) })?;
        Ok(b)
    }
    pub fn as_u32(&self) -> *mut u32 {
        self.ptr as *mut u32
    }
    pub fn download_u32(&self, words: usize) -> Result<Vec<u32>, Error> {
        let mut out = vec![0u32; words];
        self.ctx.check(unsafe { sys::b200zk_dev_download(self.ctx.raw, out.as_mut_ptr() as *mut c_void, self.ptr, (4 * words) as u64) })?;
        Ok(out)
    }
}
impl Drop for DeviceBuf {
    fn drop(&mut self) {
        if !self.ptr.is_null() {
            unsafe { sys::b200zk_dev_free(self.ctx.raw, self.ptr) }
        }
    }
}

{-# LANGUAGE ForeignFunctionInterface #-}
-- | Raw bindings to libdca_b200.so (include/dca_b200.h).  One `foreign import ccall`
-- per entry point the Haskell host modules use; nothing else.
--
-- NOT COMPILED IN THIS REPOSITORY'S CI: the build image has no GHC (the reference
-- needs GHC 8.4 / lts-12.10 + LLVM 6).  The C ABI these lines bind is exercised by
-- the ctypes host layer (haskelldca_b200/_ffi.py), whose argument lists are the same.
module DcaB200.FFI where

import Data.Int
import Data.Word
import Foreign.C.String (CString)
import Foreign.C.Types
import Foreign.Ptr

-- | opaque `dca_handle`
data DcaHandle

-- | `dca_config` (include/dca_b200.h); marshalled by DcaB200.Fused.withConfig
data DcaConfig

-- lifecycle ------------------------------------------------------------------
foreign import ccall safe "dca_config_defaults" c_dca_config_defaults :: Ptr DcaConfig -> IO CInt
foreign import ccall safe "dca_create" c_dca_create :: Ptr DcaConfig -> Ptr (Ptr DcaHandle) -> IO CInt
foreign import ccall safe "dca_destroy" c_dca_destroy :: Ptr DcaHandle -> IO CInt
foreign import ccall safe "&dca_destroy" p_dca_destroy :: FunPtr (Ptr DcaHandle -> IO CInt)
foreign import ccall safe "dca_reset" c_dca_reset :: Ptr DcaHandle -> IO CInt
foreign import ccall safe "dca_last_error" c_dca_last_error :: Ptr DcaHandle -> IO CString
foreign import ccall safe "dca_device_count" c_dca_device_count :: IO CInt
foreign import ccall safe "dca_set_rng_mt19937" c_dca_set_rng_mt19937
  :: Ptr DcaHandle -> Int32 -> Word64 -> CSize -> IO CInt

-- fused run: replaces runPeriod's loop of runStep (src/SimRunner.hs:132-178) ------
foreign import ccall safe "dca_run" c_dca_run :: Ptr DcaHandle -> Int64 -> IO CInt
-- the same in two halves / over several handles: one Haskell thread keeps a handle per GPU busy (include/dca_b200.h)
foreign import ccall safe "dca_run_async" c_dca_run_async :: Ptr DcaHandle -> Int64 -> IO CInt
foreign import ccall safe "dca_wait" c_dca_wait :: Ptr DcaHandle -> IO CInt
foreign import ccall safe "dca_run_many" c_dca_run_many :: Ptr (Ptr DcaHandle) -> Int32 -> Int64 -> IO CInt
foreign import ccall safe "dca_elapsed_ms" c_dca_elapsed_ms :: Ptr DcaHandle -> Ptr CFloat -> IO CInt

-- accFn1 / accFn2 on explicit arrays (src/SimRunner.hs:61-62, 211-212) ------------
foreign import ccall safe "dca_acc_get_action" c_dca_acc_get_action
  :: Int32 -- ^ device
  -> Int32 -- ^ order (0 = reference fold order, 1 = fast tree)
  -> Int32 -> Int32 -- ^ cell row, col
  -> Int32 -- ^ eIsEnd
  -> Ptr CFloat -- ^ wNet [3479]
  -> Ptr Word8 -- ^ grid [7*7*70] (Bool as bytes)
  -> Ptr Int64 -- ^ frep [7*7*71]
  -> Ptr Int32 -- ^ out: channel, -1 = Nothing
  -> Ptr Int64 -- ^ out: nextFrep [7*7*71]
  -> IO CInt

foreign import ccall safe "dca_acc_grid_step_train" c_dca_acc_grid_step_train
  :: Int32 -> Int32 -- ^ device, order
  -> CFloat -> CFloat -> CFloat -- ^ alphaNet, alphaAvg, alphaGrad
  -> Int32 -- ^ eIsEnd
  -> Int32 -> Int32 -- ^ cell row, col
  -> Int32 -- ^ channel, -1 = Nothing
  -> Ptr Int64 -> Ptr Int64 -- ^ frep, nextFrep
  -> Ptr Word8 -- ^ in/out: grid
  -> Ptr CFloat -> Ptr CFloat -- ^ in/out: wNet, wGradCorr
  -> Ptr CFloat -- ^ in/out: avgReward
  -> Ptr CFloat -> Ptr Int32 -- ^ out: loss, loss-is-NaN
  -> Ptr Int64 -- ^ out: reward
  -> IO CInt

-- Gridfuncs on explicit arrays ---------------------------------------------------
foreign import ccall safe "dca_gf_feature_rep" c_dca_gf_feature_rep :: Int32 -> Ptr Word8 -> Ptr Int64 -> IO CInt
foreign import ccall safe "dca_gf_violates_reuse_constraint" c_dca_gf_violates_reuse_constraint
  :: Int32 -> Ptr Word8 -> Ptr Int32 -> IO CInt

-- read-back ----------------------------------------------------------------------
foreign import ccall safe "dca_read_state" c_dca_read_state
  :: Ptr DcaHandle -> Int32 -> Ptr Word8 -> Ptr Int64 -> Ptr CFloat -> Ptr CFloat -> Ptr CFloat -> Ptr Int64
  -> Ptr DcaEvent -> IO CInt
foreign import ccall safe "dca_read_stats" c_dca_read_stats :: Ptr DcaHandle -> Int32 -> Ptr Int64 -> IO CInt
foreign import ccall safe "dca_read_status" c_dca_read_status
  :: Ptr DcaHandle -> Int32 -> Ptr Int32 -> Ptr Int64 -> Ptr Word64 -> IO CInt
foreign import ccall safe "dca_read_period" c_dca_read_period
  :: Ptr DcaHandle -> Int32 -> Ptr CFloat -> Ptr Int64 -> Int32 -> IO CInt
foreign import ccall safe "dca_read_summary" c_dca_read_summary
  :: Ptr DcaHandle -> Ptr Int32 -> Ptr Int64 -> Ptr CFloat -> Ptr Int64 -> IO CInt
foreign import ccall safe "dca_sum_stats" c_dca_sum_stats :: Ptr DcaHandle -> Ptr Int64 -> Ptr (Ptr ()) -> IO CInt
foreign import ccall safe "dca_allreduce_stats" c_dca_allreduce_stats
  :: Ptr (Ptr DcaHandle) -> Int32 -> Ptr Int64 -> IO CInt

-- | `dca_event` (include/dca_b200.h); peeked field by field in DcaB200.Fused
data DcaEvent

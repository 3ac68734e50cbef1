//! convolution(), convolution_op.rs:94-193 -> b200_conv2d (tcgen05 3xTF32 implicit GEMM, Bias / Relu fused in the epilogue).
use std::ptr;

use b200rt_sys as sys;
use onnx_protobuf::{NodeProto, TensorProto, ValueInfoProto};

use crate::device::{check, default_context, DeviceTensor, Store};
use crate::inference_engine::utils::get_stored_tensor;

pub fn convolution(output_container: &Store,
                   node: &NodeProto,
                   model_inputs: &Vec<ValueInfoProto>,
                   model_initializers: &Vec<TensorProto>) {
    let from_store = |i: usize| -> Option<DeviceTensor> {
        output_container.lock().unwrap().get(&node.input[i]).map(|v| v.1.clone().expect("Conv operands must be rank 4 (convolution_op.rs:101,111)"))
    };
    let x = from_store(0).unwrap_or_else(|| get_stored_tensor(0, node, model_inputs, model_initializers));
    let w = from_store(1).unwrap_or_else(|| get_stored_tensor(1, node, model_inputs, model_initializers));
    let bias = if node.input.len() > 2 { Some(get_stored_tensor(2, node, model_inputs, model_initializers)) } else { None };

    let mut p = sys::b200_conv_params::default();
    p.auto_pad = sys::B200_PAD_VALID;                                  // default, convolution_op.rs:134
    for attr in &node.attribute {
        match attr.name.as_ref() {
            "auto_pad" => p.auto_pad = match std::str::from_utf8(&attr.s).unwrap() {
                "SAME_UPPER" => sys::B200_PAD_SAME_UPPER,
                "SAME_LOWER" => sys::B200_PAD_SAME_LOWER,
                "VALID" => sys::B200_PAD_VALID,
                "NOT_SET" => sys::B200_PAD_NOTSET,                      // sic, convolution_op.rs:143
                other => panic!("Convolution Auto Pad specified not found: {}", other),
            },
            "dilations" => { p.dilations = [attr.ints[0], attr.ints[1]]; }
            "group" => p.group = attr.i,
            "kernel_shape" => {}                                        // taken from the weight, convolution_op.rs:151
            "pads" => { for (i, v) in attr.ints.iter().take(4).enumerate() { p.pads[i] = *v; } }
            "strides" => { p.strides = [attr.ints[0], attr.ints[1]]; }  // required: the backend rejects {0,0} like :285 unwraps
            _ => panic!("ATTRIBUTE NAME FOR CONVOLUTION NOT FOUND, {}", <String as AsRef<str>>::as_ref(&attr.name)),
        }
    }
    let ctx = default_context();
    let mut y = ptr::null_mut();
    check(unsafe {
        sys::b200_conv2d(ctx.raw(), x.raw(), w.raw(), bias.as_ref().map_or(ptr::null(), |b| b.raw() as *const _), ptr::null(), &p, &mut y)
    }).unwrap_or_else(|e| panic!("b200rt: {}", e.message));
    output_container.lock().unwrap().insert(node.output[0].clone(), (None, Some(DeviceTensor::from_raw(ctx.clone(), y))));
}

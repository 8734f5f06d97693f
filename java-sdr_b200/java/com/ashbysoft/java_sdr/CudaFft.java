// CudaFft.java — drop-in for fft.java's DSP: same constructor shape and the same
// IAudioHandler / IRawHandler / IPublishListener surface (fft.java:19,34-35,56-61,
// 190-228); the transform, PSD and peak search run in libjsdrcuda.so.
//
// NOT COMPILED HERE (no JDK in the build image); see INTEGRATION.md.  jsdr.java
// switches to it by constructing `new CudaFft(...)` where it now constructs
// `new fft(...)` (jsdr.java:476); waterfall.java keeps listening for "fft-psd".
package com.ashbysoft.java_sdr;

import java.lang.foreign.MemorySegment;
import java.lang.foreign.ValueLayout;

public class CudaFft implements IAudioHandler, IRawHandler, IPublishListener {
	private final IPublish publish;
	private final ILogger logger;
	private final JsdrCuda.Context ctx;
	private IAudio audio;
	private AudioDescriptor adsc;
	private MemorySegment handle, pinIn, pinOut, pinPeak;
	private float[] psd;                  // published array, reused every block (fft.java:68,226)
	private int n;
	private final boolean useRaw;         // true: take the s16 bytes (4 B/sample over PCIe instead of 8)

	public CudaFft(IConfig cfg, IPublish pub, ILogger lg, IUIHost hst, IAudio aud,
		JsdrCuda.Context ctx, boolean useRaw) {
		this.publish = pub;
		this.logger = lg;
		this.ctx = ctx;
		this.useRaw = useRaw;
		setup(aud);
		pub.listen(this);                                        // fft.java:53
	}

	public void notify(String key, Object val) {
		if ("audio-change".equals(key)) setup((IAudio) val);     // fft.java:56-61
	}

	private synchronized void setup(IAudio aud) {
		try {
			if (handle != null) { int rc = (int) JsdrCuda.FFT_DESTROY.invokeExact(handle); }
			audio = aud;
			adsc = audio.getAudioDescriptor();
			n = adsc.blen / adsc.size;                           // fft.java:67
			psd = new float[n + 2];                              // fft.java:68
			MemorySegment out = ctx.arena.allocate(ValueLayout.ADDRESS);
			String err = JsdrCuda.check((int) JsdrCuda.FFT_CREATE.invokeExact(ctx.handle, n, adsc.rate, 1, out));
			if (err != null) { logger.statusMsg(err); handle = null; return; }
			handle = out.get(ValueLayout.ADDRESS, 0);
			pinIn = ctx.pinned(8L * n);
			pinOut = ctx.pinned(4L * (n + 2));
			pinPeak = ctx.pinned(4);
			audio.remHandler(this);                              // fft.java:75-76
			audio.addHandler(this);
			if (useRaw) { audio.remRawHandler(this); audio.addRawHandler(this); }
		} catch (Throwable t) {
			logger.statusMsg("CudaFft setup: " + t);
			handle = null;
		}
	}

	// IAudioHandler: float IQ in +-1, 2*n floats, caller reuses buf (JavaAudio.java:224)
	public synchronized void receive(float[] buf) {
		if (handle == null || useRaw) return;
		try {
			MemorySegment.copy(buf, 0, pinIn, ValueLayout.JAVA_FLOAT, 0, 2 * n);     // fft.java:192
			String err = JsdrCuda.check((int) JsdrCuda.FFT_RECEIVE_F32.invokeExact(
				handle, pinIn, 1, pinOut, pinPeak, JsdrCuda.MEM_HOST));
			if (err != null) { logger.statusMsg(err); return; }                      // never throw (JavaAudio.java:321)
			MemorySegment.copy(pinOut, ValueLayout.JAVA_FLOAT, 0, psd, 0, n + 2);
			publish.setPublish("fft-psd", psd);                                      // fft.java:226
		} catch (Throwable t) {
			logger.statusMsg("CudaFft: " + t);
		}
	}

	// IRawHandler: blen bytes s16le IQ before I/Q correction (JavaAudio.java:262-265);
	// the correction of JavaAudio.java:281-288 is applied on the device
	public synchronized void receive(byte[] raw) {
		if (handle == null || !useRaw) return;
		try {
			MemorySegment.copy(raw, 0, pinIn, ValueLayout.JAVA_BYTE, 0, 4 * n);
			String err = JsdrCuda.check((int) JsdrCuda.FFT_RECEIVE_S16.invokeExact(
				handle, pinIn, 1, audio.getICorrection(), audio.getQCorrection(), pinOut, pinPeak, JsdrCuda.MEM_HOST));
			if (err != null) { logger.statusMsg(err); return; }
			MemorySegment.copy(pinOut, ValueLayout.JAVA_FLOAT, 0, psd, 0, n + 2);
			publish.setPublish("fft-psd", psd);
		} catch (Throwable t) {
			logger.statusMsg("CudaFft: " + t);
		}
	}
}

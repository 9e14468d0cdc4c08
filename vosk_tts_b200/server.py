"""gRPC TTS service over the CUDA engine, wire-compatible with the reference server (server/tts_server.py:35-64,
server/tts_service.proto:76-95): service `vosk.tts.Synthesizer`, server-streaming method `UtteranceSynthesis`, request
`UtteranceSynthesisRequest{model, text, hints[], output_audio_spec}`, responses `UtteranceSynthesisResponse{audio_chunk{data}}`
carrying 16-bit little-endian PCM at 22 050 Hz.  The reference's client (server/tts_client.py:21-27) works unchanged.

Differences, all on the server side of the wire:
  * the reference answers with ONE message holding the whole utterance; this server streams one message per decoder window
    (`Synth.synth_audio_stream`: text encoder, duration predictor and flow once, then the vocoder chunk by chunk), so the
    first audio leaves after ~1/10 of the total latency for long texts (`VOSK_SERVER_CHUNK_FRAMES=0` restores one message);
  * `speech_rate` is used as sent (the reference truncates it to an integer, tts_server.py:51, so 1.5 becomes 1);
  * no generated `tts_service_pb2*.py`: protoc is not needed, the seven messages are declared below with the protobuf
    runtime and the method is registered through a generic handler;
  * one `Synth` (one GPU engine handle) shared by the worker threads like in the reference (tts_server.py:37-38,57): requests
    are serialised on the engine by the session lock, g2p / PCM conversion / HTTP2 framing of different requests overlap.

Environment (same names as the reference, tts_server.py:30-33): VOSK_SERVER_INTERFACE, VOSK_SERVER_PORT, VOSK_MODEL_PATH,
VOSK_SERVER_THREADS; plus VOSK_SERVER_CHUNK_FRAMES (default 64 frames = 0.74 s of audio per message).
"""
import logging
import os
from concurrent import futures

SERVICE = "vosk.tts.Synthesizer"
METHOD = "UtteranceSynthesis"

_messages = None


def messages():
    """The message classes of server/tts_service.proto, built once from a programmatic FileDescriptorProto (field numbers,
    types, oneofs and enum values are the wire contract; there is no generated module)."""
    global _messages
    if _messages is not None:
        return _messages
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
    F = descriptor_pb2.FieldDescriptorProto
    fd = descriptor_pb2.FileDescriptorProto()
    fd.name = "vosk_tts_b200/tts_service.proto"
    fd.package = "vosk.tts"
    fd.syntax = "proto3"

    def msg(name):
        m = fd.message_type.add()
        m.name = name
        return m

    def field(m, name, number, ftype, type_name=None, label=F.LABEL_OPTIONAL, oneof=None):
        f = m.field.add()
        f.name, f.number, f.type, f.label = name, number, ftype, label
        if type_name:
            f.type_name = type_name
        if oneof is not None:
            f.oneof_index = oneof
        return f

    m = msg("AudioFormatOptions")
    m.oneof_decl.add().name = "AudioFormat"
    field(m, "raw_audio", 1, F.TYPE_MESSAGE, ".vosk.tts.RawAudio", oneof=0)
    field(m, "container_audio", 2, F.TYPE_MESSAGE, ".vosk.tts.ContainerAudio", oneof=0)

    m = msg("RawAudio")
    e = m.enum_type.add()
    e.name = "AudioEncoding"
    for n, v in (("AUDIO_ENCODING_UNSPECIFIED", 0), ("LINEAR16_PCM", 1)):
        ev = e.value.add()
        ev.name, ev.number = n, v
    field(m, "audio_encoding", 1, F.TYPE_ENUM, ".vosk.tts.RawAudio.AudioEncoding")
    field(m, "sample_rate_hertz", 2, F.TYPE_INT64)

    m = msg("ContainerAudio")
    e = m.enum_type.add()
    e.name = "ContainerAudioType"
    for n, v in (("CONTAINER_AUDIO_TYPE_UNSPECIFIED", 0), ("WAV", 1), ("OGG_OPUS", 2), ("MP3", 3)):
        ev = e.value.add()
        ev.name, ev.number = n, v
    field(m, "container_audio_type", 1, F.TYPE_ENUM, ".vosk.tts.ContainerAudio.ContainerAudioType")

    m = msg("UtteranceSynthesisResponse")
    field(m, "audio_chunk", 1, F.TYPE_MESSAGE, ".vosk.tts.AudioChunk")

    m = msg("AudioChunk")
    field(m, "data", 1, F.TYPE_BYTES)

    m = msg("Hints")
    m.oneof_decl.add().name = "Hint"
    field(m, "speaker_id", 1, F.TYPE_INT64, oneof=0)
    field(m, "speech_rate", 2, F.TYPE_DOUBLE, oneof=0)
    field(m, "role", 3, F.TYPE_STRING, oneof=0)

    m = msg("UtteranceSynthesisRequest")
    m.oneof_decl.add().name = "Utterance"
    field(m, "model", 1, F.TYPE_STRING)
    field(m, "text", 2, F.TYPE_STRING, oneof=0)
    field(m, "hints", 3, F.TYPE_MESSAGE, ".vosk.tts.Hints", label=F.LABEL_REPEATED)
    field(m, "output_audio_spec", 4, F.TYPE_MESSAGE, ".vosk.tts.AudioFormatOptions")

    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    out = {}
    for name in ("AudioFormatOptions", "RawAudio", "ContainerAudio", "UtteranceSynthesisResponse", "AudioChunk", "Hints",
                 "UtteranceSynthesisRequest"):
        out[name] = message_factory.GetMessageClass(pool.FindMessageTypeByName("vosk.tts." + name))
    _messages = out
    return out


class SynthesizerServicer:
    """`synth` is a vosk_tts_b200.Synth (or anything with synth_audio / synth_audio_stream of that signature)."""

    def __init__(self, synth, chunk_frames=64):
        self.synth = synth
        self.chunk_frames = int(chunk_frames)

    def UtteranceSynthesis(self, request, context):
        M = messages()
        speaker_id = 0
        speech_rate = 1.0
        for hint in request.hints:                       # same precedence and defaults as tts_server.py:45-52
            if hint.HasField("speaker_id"):
                speaker_id = hint.speaker_id
            if hint.HasField("speech_rate"):
                speech_rate = hint.speech_rate
        if not speech_rate > 0:
            import grpc
            context.abort(grpc.StatusCode.INVALID_ARGUMENT, "speech_rate must be positive")
        try:
            if self.chunk_frames > 0 and hasattr(self.synth, "synth_audio_stream"):
                stream = self.synth.synth_audio_stream(request.text, speaker_id=speaker_id, speech_rate=speech_rate,
                                                       chunk_frames=self.chunk_frames)
                try:
                    for pcm in stream:
                        yield M["UtteranceSynthesisResponse"](audio_chunk=M["AudioChunk"](data=pcm.tobytes()))
                finally:
                    stream.close()      # a cancelled RPC must not leave the engine locked for the next request
            else:
                audio = self.synth.synth_audio(request.text, speaker_id=speaker_id, speech_rate=speech_rate)
                yield M["UtteranceSynthesisResponse"](audio_chunk=M["AudioChunk"](data=audio.tobytes()))
        except (KeyError, ValueError) as ex:             # unknown phoneme / unsupported model type / bad speaker id
            import grpc
            context.abort(grpc.StatusCode.INVALID_ARGUMENT, "%s: %s" % (type(ex).__name__, ex))


def add_servicer(servicer, server):
    import grpc
    M = messages()
    handler = grpc.unary_stream_rpc_method_handler(servicer.UtteranceSynthesis,
                                                   request_deserializer=M["UtteranceSynthesisRequest"].FromString,
                                                   response_serializer=lambda m: m.SerializeToString())
    server.add_generic_rpc_handlers((grpc.method_handlers_generic_handler(SERVICE, {METHOD: handler}),))


def make_server(synth, address="127.0.0.1:0", threads=None, chunk_frames=64):
    """Returns (grpc server (not started), bound port)."""
    import grpc
    server = grpc.server(futures.ThreadPoolExecutor(threads or (os.cpu_count() or 1)))
    add_servicer(SynthesizerServicer(synth, chunk_frames), server)
    port = server.add_insecure_port(address)
    return server, port


def synthesize(address, text, speaker_id=None, speech_rate=None, timeout=60.0):
    """Client side of the same method (what server/tts_client.py does with the generated stub): yields the PCM bytes of
    every response message."""
    import grpc
    M = messages()
    hints = []
    if speaker_id is not None:
        hints.append(M["Hints"](speaker_id=int(speaker_id)))
    if speech_rate is not None:
        hints.append(M["Hints"](speech_rate=float(speech_rate)))
    with grpc.insecure_channel(address) as channel:
        call = channel.unary_stream("/%s/%s" % (SERVICE, METHOD), request_serializer=lambda m: m.SerializeToString(),
                                    response_deserializer=M["UtteranceSynthesisResponse"].FromString)
        for r in call(M["UtteranceSynthesisRequest"](text=text, hints=hints), timeout=timeout):
            yield r.audio_chunk.data


def serve():
    from .model import Model
    from .synth import Synth
    interface = os.environ.get("VOSK_SERVER_INTERFACE", "0.0.0.0")
    port = int(os.environ.get("VOSK_SERVER_PORT", 5001))
    model_path = os.environ.get("VOSK_MODEL_PATH", "vosk-model-tts-ru-0.8-multi")
    threads = int(os.environ.get("VOSK_SERVER_THREADS", os.cpu_count() or 1))
    chunk = int(os.environ.get("VOSK_SERVER_CHUNK_FRAMES", 64))
    synth = Synth(Model(model_path=model_path))
    server, _ = make_server(synth, "%s:%d" % (interface, port), threads, chunk)
    server.start()
    logging.info("Listening on %s:%d" % (interface, port))
    server.wait_for_termination()


if __name__ == "__main__":
    logging.basicConfig(level=logging.INFO)
    serve()

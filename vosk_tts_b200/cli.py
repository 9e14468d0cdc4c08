"""`vosk-tts` command line (flags of vosk_tts/cli.py:12-43) on top of the CUDA engine."""
import argparse
import logging
import sys

from .model import Model, list_languages, list_models
from .synth import Synth


def main(argv=None):
    p = argparse.ArgumentParser(description="Synthesize input")
    p.add_argument("--model", "-m", type=str, help="model path")
    p.add_argument("--list-models", default=False, action="store_true", help="list available models")
    p.add_argument("--list-languages", default=False, action="store_true", help="list available languages")
    p.add_argument("--model-name", "-n", type=str, help="select model by name")
    p.add_argument("--lang", "-l", default="en-us", type=str, help="select model by language")
    p.add_argument("--input", "-i", type=str, help="input string")
    p.add_argument("--speaker", "-s", type=int, help="speaker id for multispeaker model")
    p.add_argument("--speech-rate", "-r", type=float, default=1.0, help="speech rate of the synthesis")
    p.add_argument("--output", "-o", default="out.wav", type=str, help="optional output filename path")
    p.add_argument("--log-level", default="INFO", help="logging level")
    p.add_argument("--device", type=int, default=0, help="CUDA device index (extension)")
    args = p.parse_args(argv)
    logging.getLogger().setLevel(args.log_level.upper())
    if args.list_models:
        list_models()
        return 0
    if args.list_languages:
        list_languages()
        return 0
    if not args.input:
        logging.info("Please specify input text or file")
        return 1
    model = Model(args.model, args.model_name, args.lang, device=args.device)
    Synth(model).synth(args.input, args.output, speaker_id=args.speaker, speech_rate=args.speech_rate)
    return 0


if __name__ == "__main__":
    sys.exit(main())
